"""Size-independent properties of the CPU oracle (oracle/pcl_oracle.c), the checker the GPU parity
tests compare against: furthest-point sampling, ball query and kNN obey the invariants the reference
kernels' semantics imply, on seeded random clouds of several sizes."""
import numpy as np
import pytest
import torch

import oracle
from pointcloudlib_b200.synthetic import modelnet_batch


@pytest.mark.parametrize("B,N,M", [(2, 256, 64), (3, 1000, 333), (1, 4096, 512)])
def test_fps_selects_distinct_points_and_greedy_maximises_the_min_distance(B, N, M):
    xyz = modelnet_batch(B, N, seed=N)[0].numpy()
    idx = oracle.fps(xyz, M, block_size=1)
    assert idx.shape == (B, M) and (idx[:, 0] == 0).all()
    for b in range(B):
        assert len(set(idx[b].tolist())) == M                      # no point is picked twice
        # greedy property: every pick is a point at maximal distance to the picks before it
        # (points with |p|^2 <= 1e-3 are skipped by the reference, ops.py:162-163)
        p = xyz[b].astype(np.float64)
        ok = (p ** 2).sum(1) > 1e-3
        dmin = np.full(N, np.inf)
        for j in range(1, min(M, 40)):
            dmin = np.minimum(dmin, ((p - p[idx[b, j - 1]]) ** 2).sum(1))
            assert dmin[idx[b, j]] >= dmin[ok].max() * (1 - 1e-5)


@pytest.mark.parametrize("N,S,r,ns", [(512, 64, 0.15, 16), (2048, 128, 0.3, 64)])
def test_ball_query_rows_are_first_hits_in_index_order_padded_with_the_first(N, S, r, ns):
    xyz = modelnet_batch(2, N, seed=S)[0].numpy()
    new_xyz = oracle.index_points(xyz, oracle.fps(xyz, S))
    idx, cnt = oracle.ball_query(new_xyz, xyz, r, ns)
    r2 = np.float32(r) * np.float32(r)
    for b in range(2):
        d2 = ((new_xyz[b][:, None, :].astype(np.float64) - xyz[b][None].astype(np.float64)) ** 2).sum(-1)
        for s in range(S):
            c = cnt[b, s]
            row = idx[b, s]
            assert 1 <= c <= ns
            assert (np.diff(row[:c]) > 0).all()                    # ascending = index order
            assert (row[c:] == row[0]).all()                       # padding = first hit
            inside = np.flatnonzero(d2[s] < float(r2) * (1 - 1e-5))
            outside = np.flatnonzero(d2[s] > float(r2) * (1 + 1e-5))
            assert not set(row[:c].tolist()) & set(outside.tolist())
            if c < ns:                                             # nothing inside the ball is missed
                assert set(inside.tolist()) <= set(row[:c].tolist())
            else:                                                  # the first ns hits: nothing earlier is skipped
                assert set(inside[inside < row[c - 1]].tolist()) <= set(row[:c].tolist())


@pytest.mark.parametrize("C,Nq,Nr,k", [(3, 100, 300, 20), (64, 64, 128, 16)])
def test_knn_is_ascending_stable_and_complete(C, Nq, Nr, k):
    g = torch.Generator().manual_seed(C)
    x_q = torch.randn(2, C, Nq, generator=g).numpy()
    x_r = torch.randn(2, C, Nr, generator=g).numpy()
    x_r[:, :, Nr // 2:Nr // 2 + 5] = x_r[:, :, :5]                  # duplicated references -> exact ties
    idx = oracle.knn(x_q, x_r, k)                                   # (B, k, Nq)
    assert idx.shape == (2, k, Nq)
    for b in range(2):
        d = ((x_q[b].T[:, None, :].astype(np.float64) - x_r[b].T[None].astype(np.float64)) ** 2).sum(-1)
        for q in range(0, Nq, 7):
            sel = idx[b, :, q]
            assert len(set(sel.tolist())) == k
            ds = d[q, sel]
            assert (np.diff(ds) >= -1e-4 * (1 + ds[:-1])).all()     # nearest first
            assert ds.max() <= np.partition(d[q], k - 1)[k - 1] * (1 + 1e-4) + 1e-6
            for i in range(5):                                      # a duplicate pair is reported lower index first
                a, c = i, Nr // 2 + i
                if a in sel and c in sel:
                    assert list(sel).index(a) < list(sel).index(c)

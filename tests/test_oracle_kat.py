"""Known-answer tests for the CPU oracle (hand-derived from the kernel semantics).

The reference ships no golden vectors for this path, so these KATs pin the restatement to the
behaviour READ from the kernel sources; tests/test_ref_kernels_gpu.py pins it to the reference's
own compiled kernels on the GPU box.
"""
import numpy as np

import oracle


def test_optimal_block_natural_log():
    # misc/ops.py:110-111  2 ** int(math.log(B))
    assert [oracle.optimal_block(b) for b in (1, 2, 3, 8, 16, 32, 64)] == [1, 1, 2, 4, 4, 8, 16]


def test_fps_collinear_tie_rules():
    # x_i = (i+1, 0, 0), i = 0..7.  start 0 -> farthest 7 -> tie between idx 3 (x=4) and 4 (x=5)
    xyz = np.zeros((1, 8, 3), np.float32)
    xyz[0, :, 0] = np.arange(1, 9)
    # one reference thread: lowest index wins the tie
    assert oracle.fps(xyz, 4, block_size=1)[0].tolist()[:3] == [0, 7, 3]
    # 8 reference threads: tree reduce -> smallest bit-reversed tid: brev3(3)=6, brev3(4)=1 -> 4
    assert oracle.fps(xyz, 4, block_size=8)[0].tolist()[:3] == [0, 7, 4]


def test_fps_skips_near_origin_points():
    xyz = np.array([[[1, 0, 0], [0.01, 0.01, 0.01], [0, 2, 0], [0, 0, 0], [-3, 0, 0]]], np.float32)
    idx = oracle.fps(xyz, 3, block_size=1)[0].tolist()
    assert idx == [0, 4, 2]
    # all remaining candidates skipped -> besti stays 0 (ops.py:152)
    xyz2 = np.array([[[1, 0, 0], [0.01, 0, 0], [0, 0.01, 0]]], np.float32)
    assert oracle.fps(xyz2, 3, block_size=1)[0].tolist() == [0, 0, 0]


def test_ball_query_padding_strictness_and_order():
    xyz = np.array([[[0, 0, 0], [0.5, 0, 0], [1.0, 0, 0], [0.25, 0, 0], [3, 0, 0], [0.1, 0, 0]]],
                   np.float32)
    new_xyz = np.array([[[0, 0, 0], [3, 0, 0]]], np.float32)
    idx, cnt = oracle.ball_query(new_xyz, xyz, 1.0, 4)
    # d2 < r2 strictly: the point at distance exactly 1.0 is excluded; first 4 hits in index order
    assert idx[0, 0].tolist() == [0, 1, 3, 5] and cnt[0, 0] == 4
    # a single hit pads every slot with it
    assert idx[0, 1].tolist() == [4, 4, 4, 4] and cnt[0, 1] == 1
    idx, cnt = oracle.ball_query(new_xyz, xyz, 0.3, 4)
    assert idx[0, 0].tolist() == [0, 3, 5, 0] and cnt[0, 0] == 3


def test_group_layout_xyz_first_and_centred():
    xyz = np.arange(12, dtype=np.float32).reshape(1, 4, 3)
    feat = 100 + np.arange(8, dtype=np.float32).reshape(1, 4, 2)
    new_xyz = xyz[:, 1:2]
    idx = np.array([[[3, 0]]], np.int32)
    out = oracle.group(new_xyz, xyz, feat, idx, use_xyz=True)
    assert out.shape == (1, 1, 2, 5)
    np.testing.assert_array_equal(out[0, 0, 0], [6, 6, 6, 106, 107])
    np.testing.assert_array_equal(out[0, 0, 1], [-3, -3, -3, 100, 101])
    out = oracle.group(new_xyz, xyz, feat, idx, use_xyz=False)
    np.testing.assert_array_equal(out[0, 0, 0], [106, 107])


def test_knn_duplicates_lower_index_first_and_kmajor_layout():
    # refs (B=1, C=1, Nr=5): values 0, 1, 1, 2, 0 ; queries: 1.0 and 0.0
    x_r = np.array([[[0, 1, 1, 2, 0]]], np.float32)
    x_q = np.array([[[1.0, 0.0]]], np.float32)
    idx = oracle.knn(x_q, x_r, 3)
    assert idx.shape == (1, 3, 2)
    assert idx[0, :, 0].tolist() == [1, 2, 0]  # dist 0,0 (1 before 2), then dist 1: idx 0 before 3, 4
    assert idx[0, :, 1].tolist() == [0, 4, 1]


def test_three_nn_stable_and_weights():
    xyz2 = np.array([[[0, 0, 0], [1, 0, 0], [1, 0, 0], [5, 0, 0]]], np.float32)
    xyz1 = np.array([[[1, 0, 0]]], np.float32)
    idx, dist, w = oracle.three_nn(xyz1, xyz2)
    assert idx[0, 0].tolist() == [1, 2, 0]
    np.testing.assert_allclose(dist[0, 0], [0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(w[0, 0].sum(), 1.0, rtol=1e-6)
    assert w[0, 0, 0] == w[0, 0, 1] and w[0, 0, 2] < 1e-7


def test_knn_point_matches_numpy_argsort():
    rng = np.random.default_rng(0)
    xyz = rng.standard_normal((2, 64, 3)).astype(np.float32)
    new_xyz = xyz[:, :16]
    idx, dist = oracle.knn_point(8, xyz, new_xyz, return_dist=True)
    d = oracle.square_distance(new_xyz, xyz)
    ref = np.argsort(d, axis=-1, kind="stable")[:, :, :8]
    np.testing.assert_array_equal(idx, ref)
    np.testing.assert_array_equal(dist, np.take_along_axis(d, ref, -1))


def test_fps_pointconv_first_max_and_start():
    xyz = np.zeros((1, 8, 3), np.float32)
    xyz[0, :, 0] = np.arange(1, 9)
    assert oracle.fps_pointconv(xyz, 3, np.array([0], np.int32))[0].tolist() == [0, 7, 3]
    assert oracle.fps_pointconv(xyz, 2, np.array([7], np.int32))[0].tolist() == [7, 0]


def test_density_against_numpy():
    rng = np.random.default_rng(1)
    xyz = rng.standard_normal((2, 50, 3)).astype(np.float32)
    out = oracle.compute_density(xyz, 0.5)
    d = ((xyz[:, :, None] - xyz[:, None]) ** 2).sum(-1).astype(np.float64)
    ref = (np.exp(-d / (2 * 0.25)) / (2.5 * 0.5)).mean(-1)
    np.testing.assert_allclose(out, ref, rtol=1e-4)

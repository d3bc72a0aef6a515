"""Pin the C oracle to outputs of the REFERENCE'S OWN kernels without a GPU.

tests/golden/ref_kernels_b200.npz was written by tests/golden/make_golden.py on a B200: the kernel
strings of /root/reference/misc/ops.py (FPS :124-234, ball query :291-330, KNN :429-552), compiled
unmodified (oracle/build_ref.py) and launched with the reference's own configuration, on small seeded
inputs that are stored in the file.  The CPU restatement (oracle/pcl_oracle.c) must reproduce every
index bit-exactly; the same comparison runs against the live kernels in tests/test_ref_kernels_gpu.py."""
import os

import numpy as np
import pytest

import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_kernels_b200.npz"))


@pytest.mark.parametrize("tag", ["fps_c1", "fps_b32", "fps_adv"])
def test_oracle_fps_equals_reference_kernel_output(tag):
    xyz, idx, bs = G[tag + "_xyz"], G[tag + "_idx"], int(G[tag + "_bs"])
    assert bs == oracle.optimal_block(xyz.shape[0])
    np.testing.assert_array_equal(oracle.fps(xyz, idx.shape[1], block_size=bs), idx)
    assert (idx[:, 0] == 0).all()                                   # ops.py:143-144: starts at index 0


@pytest.mark.parametrize("tag", ["bq_pad", "bq_full", "bq_adv"])
def test_oracle_ball_query_equals_reference_kernel_output(tag):
    xyz, new_xyz, r = G[tag + "_xyz"], G[tag + "_new_xyz"], float(G[tag + "_r"])
    idx, cnt = G[tag + "_idx"], G[tag + "_cnt"]
    oidx, ocnt = oracle.ball_query(new_xyz, xyz, r, idx.shape[2])
    np.testing.assert_array_equal(ocnt, cnt)
    np.testing.assert_array_equal(oidx, idx)
    if tag == "bq_pad":                                             # some rows are padded with the first hit
        short = cnt < idx.shape[2]
        assert short.any() and (idx[short][:, -1] == idx[short][:, 0]).all()


@pytest.mark.parametrize("tag", ["knn_xyz", "knn_feat", "knn_dup"])
def test_oracle_knn_equals_reference_kernel_output(tag):
    x_q, x_r, idx = G[tag + "_q"], G[tag + "_r"], G[tag + "_idx"]
    np.testing.assert_array_equal(oracle.knn(x_q, x_r, idx.shape[1]), idx)
    if tag == "knn_dup":                                            # equal distances: lower reference index first
        assert (idx[:, 0, :24] == np.arange(24)).all()

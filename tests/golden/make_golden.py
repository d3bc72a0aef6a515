"""Generate tests/golden/ref_kernels_b200.npz: outputs of the REFERENCE'S OWN CUDA kernels
(oracle/_ref/libref_kernels.so = the kernel strings of /root/reference/misc/ops.py compiled
unmodified by oracle/build_ref.py, launched with the reference's grid = B, block = optimal_block(B))
on small seeded inputs, run on a B200.  The inputs are stored next to the outputs, so the CPU-only
suite (tests/test_golden_cpu.py) pins the C oracle to the reference without a GPU.

    gpurun -- 'python tests/golden/make_golden.py'      # writes gpurun_out/ref_kernels_b200.npz
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402
from pointcloudlib_b200.synthetic import adversarial_cloud, modelnet_batch  # noqa: E402


def main():
    so = build_ref.build()
    lib = ctypes.CDLL(so)
    P, I, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.ref_fps.argtypes = [P, I, I, I, I, P, P, P]
    lib.ref_ball_query.argtypes = [P, P, I, I, I, Fl, I, I, P, P, P]
    lib.ref_knn.argtypes = [P, P, I, I, I, I, I, P, P]
    st = torch.cuda.current_stream().cuda_stream
    out = {}

    # FPS: BASELINE config 1 (B=2, N=1024, M=512) + a block-size-8 case (B=32: bit-reversed tie rule) + adversarial
    for tag, xyz, M in (("fps_c1", modelnet_batch(2, 1024, seed=0)[0], 512),
                        ("fps_b32", modelnet_batch(32, 256, seed=1)[0], 64),
                        ("fps_adv", adversarial_cloud(8, 512, seed=1), 128)):
        B, N, _ = xyz.shape
        bs = oracle.optimal_block(B)
        xd = xyz.cuda()
        temp = torch.empty(B, N, device="cuda")
        idx = torch.empty(B, M, dtype=torch.int32, device="cuda")
        assert lib.ref_fps(xd.data_ptr(), B, N, M, bs, temp.data_ptr(), idx.data_ptr(), st) == 0
        torch.cuda.synchronize()
        out[tag + "_xyz"], out[tag + "_idx"], out[tag + "_bs"] = xyz.numpy(), idx.cpu().numpy(), np.int32(bs)

    # ball query: padding (few hits), early exit (many hits), adversarial duplicates
    for tag, xyz, S, r, ns in (("bq_pad", modelnet_batch(2, 1024, seed=2)[0], 64, 0.1, 32),
                               ("bq_full", modelnet_batch(2, 1024, seed=3)[0], 64, 0.4, 16),
                               ("bq_adv", adversarial_cloud(4, 512, seed=2), 64, 0.2, 32)):
        B, N, _ = xyz.shape
        fidx = oracle.fps(xyz.numpy(), S)
        new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
        idx = torch.zeros(B, S, ns, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(B, S, dtype=torch.int32, device="cuda")
        r32 = float(str(r))
        nd, xd = new_xyz.cuda(), xyz.cuda()          # keep both alive across the launch
        assert lib.ref_ball_query(nd.data_ptr(), xd.data_ptr(), B, N, S, r32, ns,
                                  oracle.optimal_block(B), idx.data_ptr(), cnt.data_ptr(), st) == 0
        torch.cuda.synchronize()
        out[tag + "_xyz"], out[tag + "_new_xyz"] = xyz.numpy(), new_xyz.numpy()
        out[tag + "_r"], out[tag + "_idx"], out[tag + "_cnt"] = np.float64(r32), idx.cpu().numpy(), cnt.cpu().numpy()

    # KNN: xyz (C=3), feature-space (C=64), duplicates (tie order), ragged sizes
    g = torch.Generator().manual_seed(7)
    dup = torch.randn(2, 5, 40, generator=g)
    dup = torch.cat([dup, dup[:, :, :24]], dim=2)                 # duplicated references
    for tag, x_q, x_r, k in (("knn_xyz", None, modelnet_batch(2, 256, seed=4)[0].permute(0, 2, 1).contiguous(), 20),
                             ("knn_feat", torch.randn(2, 64, 96, generator=g), torch.randn(2, 64, 128, generator=g), 16),
                             ("knn_dup", dup[:, :, :50].contiguous(), dup, 12)):
        if x_q is None:
            x_q = x_r.clone()
        B, C, Nq = x_q.shape
        Nr = x_r.shape[2]
        qd, rd = x_q.cuda(), x_r.cuda()
        tmp = torch.empty(B, Nr, Nq, device="cuda")
        idx = torch.empty(B, k, Nq, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        lib.ref_knn(rd.data_ptr(), qd.data_ptr(), B, C, Nr, Nq, k, tmp.data_ptr(), idx.data_ptr())
        torch.cuda.synchronize()
        out[tag + "_q"], out[tag + "_r"], out[tag + "_idx"] = x_q.numpy(), x_r.numpy(), idx.cpu().numpy()

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_kernels_b200.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main()

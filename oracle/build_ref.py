"""Compile the reference's OWN inline CUDA kernels standalone -> oracle/_ref/libref_kernels.so.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference keeps its three custom kernels as Python string literals handed to ``jt.code``
(misc/ops.py:115-252 FPS, :290-338 ball query, :425-649 KNN).  Jittor is not installable here,
but the strings are plain CUDA C.  This script reads them from /root/reference at build time
(never copied into the repo: the generated .cu lives in a temp dir and only the compiled .so is
kept, under the git-ignored oracle/_ref/), wraps each in an ``extern "C"`` launcher that
reproduces the reference's launch configuration (grid = B, block = optimal_block(B), dynamic smem
2*block*4 for FPS; the KNN string's own knn_cuda_global host function is called as is), and
compiles for sm_100a.  The .so travels to the GPU box with the repo snapshot; when
/root/reference is absent (the GPU box) this script is a no-op and tests use the prebuilt file.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_OPS = "/root/reference/misc/ops.py"
OUT_DIR = os.path.join(_HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "libref_kernels.so")

_LAUNCHERS = r'''
extern "C" int ref_fps(const float* xyz, int B, int N, int M, int block_size, float* temp,
                       int* idx, void* stream) {
    furthest_point_sampling_kernel<<<B, block_size, 2 * block_size * sizeof(int),
                                     (cudaStream_t)stream>>>(B, N, M, block_size, xyz, temp, idx);
    return (int)cudaGetLastError();
}
extern "C" int ref_ball_query(const float* new_xyz, const float* xyz, int B, int N, int S,
                              float radius, int nsample, int block_size, int* idx, int* cnt,
                              void* stream) {
    query_ball_point_kernel<<<B, block_size, 0, (cudaStream_t)stream>>>(
        B, N, S, radius, nsample, new_xyz, xyz, idx, cnt);
    return (int)cudaGetLastError();
}
/* x_r (B,C,Nr), x_q (B,C,Nq), tmp_dist (B,Nr,Nq), idx (B,k,Nq); default stream as the reference */
extern "C" int ref_knn(float* x_r, float* x_q, int B, int C, int Nr, int Nq, int k,
                       float* tmp_dist, int* idx) {
    knn_cuda_global(B, x_r, Nr, x_q, Nq, C, k, idx, tmp_dist);
    return (int)cudaGetLastError();
}
'''


def _extract(src: str):
    fps = re.search(r"class FurthestPointSampler.*?cuda_src='''(.*?)'''", src, re.S).group(1)
    fps = fps.split("int block_size = #block_size;")[0]
    bq = re.search(r"class BallQueryGrouper.*?cuda_src = '''(.*?)'''", src, re.S).group(1)
    bq = bq.split("int block_size = #block_size;")[0]
    knn = re.search(r'self\.cuda_inc= """(.*?)"""', src, re.S).group(1)
    knn = knn.replace('#include "helper_cuda.h"', "").replace("#undef out", "")
    knn = knn.replace("\\\\n", "\\n")
    return fps, bq, knn


def build(force: bool = False) -> str | None:
    """Returns the .so path, or None when neither the reference tree nor a prebuilt .so exists."""
    if not os.path.exists(REF_OPS):
        return OUT_SO if os.path.exists(OUT_SO) else None
    if (not force) and os.path.exists(OUT_SO) and \
            os.path.getmtime(OUT_SO) >= max(os.path.getmtime(REF_OPS), os.path.getmtime(__file__)):
        return OUT_SO
    fps, bq, knn = _extract(open(REF_OPS).read())
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        cu = os.path.join(td, "ref_kernels.cu")
        with open(cu, "w") as f:
            f.write("#include <cuda_runtime.h>\n#include <cstdio>\n")
            f.write(fps + "\n" + bq + "\n" + knn + "\n" + _LAUNCHERS)
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-w", "-shared",
               "-Xcompiler", "-fPIC", "-o", OUT_SO, cu]
        subprocess.run(cmd, check=True)
    return OUT_SO


REF_ROOT = "/root/reference"
SNAPSHOT_DIR = os.path.join(os.path.dirname(_HERE), "baseline", "_ref", "PointCloudLib")
_SNAPSHOT_FILES = ("misc/layers.py", "misc/ops.py", "misc/pointconv_utils.py", "misc/utils.py",
                   "networks/cls/pointnet.py", "networks/cls/pointnet2.py", "networks/cls/dgcnn.py",
                   "networks/cls/pointconv.py", "networks/cls/pointcnn.py",
                   "networks/seg/pointnet_partseg.py", "networks/seg/pointnet2_partseg.py",
                   "networks/seg/dgcnn_partseg.py", "networks/seg/pointconv_partseg.py",
                   "networks/seg/pointcnn_partseg.py")


def snapshot() -> str | None:
    """Place the UNMODIFIED reference network / misc files (a dozen .py files) under the git-ignored
    baseline/_ref/PointCloudLib/ so they travel to the GPU box with the repo snapshot: the GPU tests
    import them from there through compat/ (``PCL_REFERENCE``) and check that the reference's own
    ``networks/**`` run on libpcl_b200.  Nothing is committed; a no-op where /root/reference is absent."""
    import shutil
    if not os.path.isdir(os.path.join(REF_ROOT, "networks")):
        return SNAPSHOT_DIR if os.path.isdir(os.path.join(SNAPSHOT_DIR, "networks")) else None
    for rel in _SNAPSHOT_FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(SNAPSHOT_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
    return SNAPSHOT_DIR


def reference_checkout() -> str | None:
    """Where the reference's python files can be imported from: $PCL_REFERENCE, the build container's
    /root/reference, or the snapshot made by snapshot()."""
    for cand in (os.environ.get("PCL_REFERENCE"), REF_ROOT, SNAPSHOT_DIR):
        if cand and os.path.isdir(os.path.join(cand, "networks")):
            return cand
    return None


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(snapshot())

"""pointcloudlib_b200 — B200-native set-abstraction / EdgeConv operators behind the reference's
``misc/ops.py`` / ``misc/pointconv_utils.py`` module API.

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/pcl_b200.h`` -> ``libpcl_b200.so``),
``_lib.py`` ctypes binding, ``functional.py`` torch-tensor front end, ``misc/`` and ``networks/``
host-side mirrors of the reference interface, ``train.py`` step plumbing (flat bucket, NCCL DP).
"""
import torch as _torch

# The reference computes in fp32.  Library GEMM/conv calls that remain on the path (dense heads,
# not-yet-fused layers) must not silently drop to TF32; the hand-written tensor-core kernels
# choose their own arithmetic explicitly (3xTF32 split, fp32-equivalent).
_torch.backends.cudnn.allow_tf32 = False
_torch.backends.cuda.matmul.allow_tf32 = False

__version__ = "0.1.0"

"""The C-ABI library loads and exports every symbol include/pcl_b200.h declares (no compute)."""
import ctypes
import os

import pytest

from pointcloudlib_b200 import _lib


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from pointcloudlib_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in pcl_b200.h but not exported: {missing}"
    # and the python binding table covers exactly the header
    bound = set(_lib._SIGNATURES) | set(_lib.FUSED_SYMBOLS) | {"pcl_last_error"}
    assert bound == set(declared), (sorted(bound - set(declared)), sorted(set(declared) - bound))


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    assert lib.pcl_compiled_arch() == 100
    # null pointers / bad shapes are rejected before any launch
    assert lib.pcl_fps(None, 1, 8, 4, 1, None, None) == -1  # non-empty problem, null buffers
    assert b"null" in lib.pcl_last_error()
    assert lib.pcl_optimal_block(32) == 8 and lib.pcl_optimal_block(2) == 1
    assert lib.pcl_optimal_block(16) == 4


def test_cpu_tensor_fails_loudly():
    import torch
    from pointcloudlib_b200 import functional as F
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.furthest_point_sample(torch.zeros(1, 8, 3), 4)


def test_routed_last_layer_backward_is_validated_before_any_launch():
    """PCL_EPI_BWD_Y_MASK_ROUTED / pcl_routed_sort reject what the kernels do not cover (no GPU needed: the checks run
    on the host and return PCL_ERR_UNSUPPORTED / PCL_ERR_INVALID before a launch)."""
    from pointcloudlib_b200 import fused
    fused._bind()
    lib = _lib.lib()._cdll
    a = fused.PclRowGemm()
    dummy = ctypes.create_string_buffer(4096)
    p = ctypes.addressof(dummy)
    for f in ("W", "x0", "x1", "selpos", "scale", "shift", "out", "stats", "ebias"):
        setattr(a, f, p)
    a.P, a.K, a.N, a.ldw, a.ns, a.C3 = 4096, 96, 96, 224, 128, 128
    # without the warp-specialised tcgen05 core (x3 == 3) the epilogue does not exist
    assert lib.pcl_rowgemm(ctypes.byref(a), fused.PRO_BN_ACT, fused.EPI_BWD_Y_MASK_ROUTED, 2, None) == -2
    assert b"PCL_EPI_BWD_Y_MASK_ROUTED" in lib.pcl_last_error()
    # K != N (the ReLU mask is the sign of operand (p, n)), C3 not a multiple of 32, a tile's entry list too long
    for field, value in (("K", 64), ("C3", 48), ("ns", 16)):
        b = fused.PclRowGemm.from_buffer_copy(a)
        setattr(b, field, value)
        if field == "ns":
            b.C3 = 128                                        # (128 / 16) * 128 = 1024 entries per tile > 512
        assert lib.pcl_rowgemm(ctypes.byref(b), fused.PRO_BN_ACT, fused.EPI_BWD_Y_MASK_ROUTED, 3, None) == -2, field
    assert lib.pcl_routed_sort(None, None, 4, 128, 128, 96, None, None) == -1
    assert lib.pcl_routed_sort(p, p, 4, 128, 96, 96, p, None) == -1     # ns must be a power of two <= 128

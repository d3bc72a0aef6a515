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

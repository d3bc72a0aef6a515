"""ctypes binding of libpcl_b200.so (the C ABI declared in include/pcl_b200.h).

The product path has NO fallback: if the shared library is missing or a tensor is not on a CUDA
device, calls raise.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcl_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pcl_b200.h")

_lib = None

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
c_longlong = ctypes.c_longlong
P = c_void_p

# name -> argtypes (all return int except pcl_last_error)
_SIGNATURES = {
    "pcl_version": [],
    "pcl_compiled_arch": [],
    "pcl_optimal_block": [c_int],
    "pcl_fps": [P, c_int, c_int, c_int, c_int, P, P],
    "pcl_gather_xyz": [P, P, c_int, c_int, c_int, P, P],
    "pcl_fps_pointconv": [P, c_int, c_int, c_int, P, P, P],
    "pcl_ball_query": [P, P, c_int, c_int, c_int, c_float, c_int, P, P, P],
    "pcl_group": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P],
    "pcl_ball_query_group": [P, P, P, c_int, c_int, c_int, c_float, c_int, c_int, c_int, P, P, P, P],
    "pcl_group_backward": [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P],
    "pcl_ball_query_msg": [P, P, c_int, c_int, c_int, c_int, P, P, P, P, P],
    "pcl_ball_query_group_msg": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P],
    "pcl_index_points": [P, P, c_int, c_int, c_int, c_int, P, P],
    "pcl_index_points_backward": [P, P, c_int, c_int, c_int, c_int, P, P],
    "pcl_knn": [P, P, c_int, c_int, c_int, c_int, c_int, P, P],
    "pcl_square_distance": [P, P, c_int, c_int, c_int, c_int, P, P],
    "pcl_knn_point": [c_int, P, P, c_int, c_int, c_int, c_int, P, P, P],
    "pcl_three_nn": [P, P, c_int, c_int, c_int, P, P, P, P],
    "pcl_three_interpolate": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "pcl_three_interpolate_backward": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "pcl_graph_feature": [P, P, c_int, c_int, c_int, c_int, P, P],
    "pcl_graph_feature_backward": [P, P, c_int, c_int, c_int, c_int, P, P],
    "pcl_compute_density": [P, c_int, c_int, c_float, P, P],
    "pcl_density_contract": [P, P, P, c_longlong, c_longlong, c_longlong, c_longlong, c_int, c_int, c_int, c_int, c_int,
                             P, P],
    "pcl_density_contract_backward": [P, P, P, P, c_longlong, c_longlong, c_longlong, c_longlong, c_int, c_int, c_int,
                                      c_int, c_int, P, P, P, P],
    "pcl_sgd_momentum": [P, P, P, c_size_t, c_float, c_float, c_float, c_float, P],
    "pcl_pack_weight": [P, c_int, c_int, c_int, c_float, P, P],
}


# entry points bound in pointcloudlib_b200/fused.py (struct-taking signatures)
FUSED_SYMBOLS = ("pcl_rowgemm", "pcl_wgrad", "pcl_gather_stats", "pcl_bn_param",
                 "pcl_maxpool_finalize", "pcl_maxpool_backward", "pcl_sel_outer",
                 "pcl_gather_bn_backward", "pcl_gather_maxmin", "pcl_gather_bn_backward_routed",
                 "pcl_gather_bn_backward_masked", "pcl_bn_act_forward", "pcl_bn_act_backward", "pcl_bn_bwd_apply",
                 "pcl_sa_bwd_prepare", "pcl_sa_bwd_finish", "pcl_sa_bwd_sums1", "pcl_routed_sort",
                 "pcl_sel_outer_sorted")


def declared_symbols(header: str = HEADER_PATH):
    """Every function name include/pcl_b200.h declares."""
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcl_[a-z0-9_]+)\s*\(", src)))


def lib() -> ctypes.CDLL:
    """Load libpcl_b200.so; raises if it has not been built (python -m pointcloudlib_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is not built "
                "(run `python -m pointcloudlib_b200.build`); there is no CPU fallback")
        l = ctypes.CDLL(LIB_PATH)
        l.pcl_last_error.restype = ctypes.c_char_p
        l.pcl_last_error.argtypes = []
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = c_int
            fn.argtypes = argtypes
        _lib = _TimedLib(l)
    return _lib


class _TimedLib:
    """Proxy over the CDLL: when a KernelTimer is active every pcl_* call is bracketed by a CUDA
    event pair on the current stream (key = the call's integer arguments)."""

    def __init__(self, cdll):
        self._cdll = cdll
        self._cache = {}

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if not name.startswith("pcl_") or name in ("pcl_last_error", "pcl_version",
                                                    "pcl_compiled_arch", "pcl_optimal_block"):
            return fn
        w = self._cache.get(name)
        if w is None:
            def w(*args, _fn=fn, _name=name):
                if not _in_call:
                    LAUNCH_TAGS[_name] += 1
                t = _timer
                if t is None or _in_call or (t.only is not None and _name not in t.only):
                    return _fn(*args)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = _fn(*args)
                e1.record()
                t.records.append((_name, tuple(a for a in args if isinstance(a, int) and abs(a) < (1 << 31)),
                                  e0, e1))
                return rc
            self._cache[name] = w
        return w


_in_call = False  # set while call() itself is timing (avoids double records)


LAUNCHES = 0          # C-ABI calls that enqueued a kernel (bench.py reads the delta)
LAUNCH_TAGS = __import__("collections").Counter()   # per entry point and per call-site tag ("sa_l3", ...)
_timer = None         # active KernelTimer or None


class KernelTimer:
    """Records a CUDA-event pair around selected C-ABI calls (on the launching stream)."""

    def __init__(self, only=None):
        self.only = set(only) if only else None
        self.records = []  # (name, key, ev_start, ev_end)

    def __enter__(self):
        global _timer
        _timer = self
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None

    def summary(self):
        """{(name, key): (n_launches, mean_ms, total_ms)} — call after a synchronize."""
        out = {}
        for name, key, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            n, tot = out.get((name, key), (0, 0.0))
            out[(name, key)] = (n + 1, tot + ms)
        return {k: (n, tot / n, tot) for k, (n, tot) in out.items()}


def call(name: str, *args, key=None) -> None:
    """Invoke a C-ABI entry point, raise on a non-zero return, time it if a KernelTimer is active."""
    global _in_call
    fn = getattr(lib(), name)
    LAUNCH_TAGS[name] += 1
    if key and isinstance(key[0], str):
        LAUNCH_TAGS[key[0]] += 1
    t = _timer
    timed = t is not None and (t.only is None or name in t.only)
    if timed:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
    _in_call = True
    try:
        if timed:
            e0.record()
        rc = fn(*args)
        if timed:
            e1.record()
    finally:
        _in_call = False
    if timed:
        t.records.append((name, key, e0, e1))
    check(rc, name)


_DEBUG_SYNC = os.environ.get("PCL_DEBUG_SYNC", "0") == "1"  # debugging aid: sync after every call


def check(rc: int, what: str = "") -> None:
    global LAUNCHES
    LAUNCHES += 1
    if _DEBUG_SYNC:
        torch.cuda.synchronize()
    if rc != 0:
        msg = lib().pcl_last_error().decode(errors="replace")
        kind = "invalid argument" if rc == -1 else "unsupported" if rc == -2 else f"cuda error {rc}" if rc > 0 else f"error {rc}"
        raise RuntimeError(f"libpcl_b200 {what}: {kind}: {msg}")


# Host-fed inputs of a CUDA-graph capture: values that the reference draws on the HOST every step (PointConv's
# random FPS start indices, np.random.randint at misc/pointconv_utils.py:88).  Code that runs under capture stages
# them through a pinned buffer and registers a refill callback here; the trainer that captured the graph calls the
# callbacks before every replay, so a replayed step draws fresh values exactly as an eager step does.
GRAPH_HOST_FEEDS = []


def host_feed(make_values, device):
    """make_values() -> 1-D int32/float32 CPU tensor.  Eager: plain copy.  Under capture: pinned staging buffer +
    a captured async copy, refilled by the registered callback before each replay."""
    vals = make_values()
    if not (device.type == "cuda" and torch.cuda.is_current_stream_capturing()):
        return vals.to(device)
    pinned = torch.empty_like(vals).pin_memory()
    pinned.copy_(vals)
    GRAPH_HOST_FEEDS.append(lambda: pinned.copy_(make_values()))
    return pinned.to(device, non_blocking=True)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libpcl_b200 operators need CUDA tensors: there is no CPU fallback "
                           f"(got a tensor on {t.device})")
    if not t.is_contiguous():
        raise RuntimeError("libpcl_b200 operators need contiguous tensors")
    return t.data_ptr()


def stream(t=None):
    dev = t.device if t is not None else None
    return torch.cuda.current_stream(dev).cuda_stream


def f32(t):
    """contiguous fp32 view/copy of a CUDA tensor."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def i32(t):
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()

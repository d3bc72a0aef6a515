"""misc/layers.py as the reference's networks import it.

When a reference checkout is reachable ($PCL_REFERENCE, /root/reference, or the git-ignored snapshot
baseline/_ref/PointCloudLib made by ``__graft_entry__.build()``), this module IS the reference's own
misc/layers.py, executed unmodified on the jittor shim — its ``from misc.ops import
FurthestPointSampler, KNN`` resolves to compat/misc/ops (libpcl_b200) — with ONE override:
``PointCNN.select_region`` (layers.py:381-388, a per-sample Python loop + jt.stack) becomes a single
gather kernel.  Without a checkout only the dense building blocks the vanilla PointNet models need
(T-Nets, Dense_Conv*, ...) are served, from pointcloudlib_b200.misc.layers.
"""
import os as _os

from pointcloudlib_b200.misc import layers as _src

from ._bridge import export as _export

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))


def _reference_layers_file():
    for cand in (_os.environ.get("PCL_REFERENCE"), "/root/reference",
                 _os.path.join(_ROOT, "baseline", "_ref", "PointCloudLib")):
        if cand and _os.path.isfile(_os.path.join(cand, "misc", "layers.py")):
            return _os.path.join(cand, "misc", "layers.py")
    return None


REFERENCE_FILE = _reference_layers_file()
if REFERENCE_FILE is not None:
    with open(REFERENCE_FILE) as _f:
        exec(compile(_f.read(), REFERENCE_FILE, "exec"), globals())   # noqa: S102 — the unmodified reference file

    def _select_region(self, pts, pts_idx):
        from jittor import _wrap
        return _wrap(_src.select_region(pts, pts_idx))

    PointCNN.select_region = _select_region   # noqa: F821 — defined by the exec above
else:
    _export(_src, globals())

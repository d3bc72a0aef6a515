"""misc/ops.py of the reference, served by pointcloudlib_b200.misc.ops (same names and signatures);
module outputs are jittor-compat Vars so the reference's network files can keep calling
``.transpose(0,3,1,2)``, ``.argmax(dim)[1]`` etc. on them."""
from pointcloudlib_b200.misc import ops as _src

from ._bridge import export as _export

_export(_src, globals())


from jittor import _wrap as _wrap_var  # noqa: E402
from pointcloudlib_b200 import lazy as _lazy  # noqa: E402

_EagerBallQueryGrouper = BallQueryGrouper  # noqa: F821 — exported above


class BallQueryGrouper(_EagerBallQueryGrouper):
    """misc/ops.py:289-407 with a DEFERRED result (pointcloudlib_b200.lazy.LazyGrouped): the reference's
    `grouper -> transpose -> mlps -> transpose -> argmax(dim=2)[1]` (networks/cls/pointnet2.py:51-57)
    then runs on the fused kernels and the (B,S,ns,3+C) tensor is never written; any other use of the
    result materialises exactly what the reference's grouper returns."""

    def __call__(self, new_xyz, pointset, feature):
        handle = _lazy.defer(self, new_xyz, pointset, feature, wrap=_wrap_var)
        if handle is not None:
            return handle
        return _EagerBallQueryGrouper.__call__(self, new_xyz, pointset, feature)

"""misc/pointconv_utils.py of the reference, served by pointcloudlib_b200.misc.pointconv_utils (same names and signatures);
module outputs are jittor-compat Vars so the reference's network files can keep calling
``.transpose(0,3,1,2)``, ``.argmax(dim)[1]`` etc. on them."""
from pointcloudlib_b200.misc import pointconv_utils as _src

from ._bridge import export as _export

_export(_src, globals())

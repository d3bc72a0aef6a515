"""`misc` as the reference's networks import it (`from misc.ops import ...`), served by
pointcloudlib_b200.misc: the reference's own misc/*.py is the path libpcl_b200 replaces."""

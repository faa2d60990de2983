"""Import the UNMODIFIED reference (graphtools v2.1.0) in the build container.

Test infrastructure only.  The reference lives at ``$GT_REF_PATH`` or
``/root/reference`` (read-only, absent on the GPU box); its three uninstalled
dependencies (`tasklogger`, `future`, `pygsp`) are replaced by the stand-ins under
``oracle/shims``.  ``NUMBA_AVAILABLE`` is forced off so the float64 path is used
(the reference's own exact-value tests do the same,
test/test_knn.py:263-264).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_path():
    for cand in (os.environ.get("GT_REF_PATH"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "graphtools")):
            return cand
    return None


def load_reference():
    """Returns the reference ``graphtools`` module, or None when it is not on this box."""
    path = reference_path()
    if path is None:
        return None
    shims = os.path.join(_HERE, "shims")
    for p in (shims, path):
        if p not in sys.path:
            sys.path.insert(0, p)
    import graphtools
    import graphtools.graphs

    graphtools.graphs.NUMBA_AVAILABLE = False
    return graphtools

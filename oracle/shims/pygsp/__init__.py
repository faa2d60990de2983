"""Stand-in for `pygsp` (test infrastructure only; reference base.py:17, 991-1008)."""
from . import graphs, utils  # noqa: F401

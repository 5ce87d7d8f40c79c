"""`import nlopt` for programs written against the reference's Python module (src/swig/nlopt-python.i): put this
directory on PYTHONPATH and the same script runs on libnlopt_b200.so.  Everything is re-exported from nlopt_b200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlopt_b200 as _m  # noqa: E402
from nlopt_b200 import *  # noqa: E402,F401,F403

for _k in dir(_m):
    if _k.isupper() or _k in ("opt", "algorithm_name", "version_major", "version_minor", "version_bugfix", "RoundoffLimited", "ForcedStop"):
        globals()[_k] = getattr(_m, _k)
roundoff_limited, forced_stop = _m.RoundoffLimited, _m.ForcedStop

"""`utils` package shim: B200-backed modules first, the reference's remaining utils after."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_ref = os.environ.get("ATOMNAS_REFERENCE")
_candidates = [os.path.join(_ref, "utils")] if _ref else []
for _p in sys.path:
    _u = os.path.join(_p, "utils")
    if os.path.isdir(_u) and os.path.abspath(_u) != _here and os.path.exists(os.path.join(_u, "config.py")):
        _candidates.append(_u)
for _u in _candidates:
    if _u not in __path__:
        __path__.append(_u)

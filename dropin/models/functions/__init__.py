"""Loss modules (losses.py, vnl.py) are outside the hot path: let them resolve to the reference checkout."""
import os
import sys

for _p in sys.path:
    _cand = os.path.join(_p, "models", "functions")
    if os.path.isfile(os.path.join(_cand, "losses.py")) and _cand not in __path__:
        __path__.append(_cand)

"""Drop-in for the reference's `models.functions.funcs`: the helpers outside the hot path (image resizing, PCA, IoU utilities
used by eval.py / simple_inference.py) are the reference's own, loaded from its checkout on sys.path; `bias_init_with_prob`
(planerecnet.py:138-141) is this repo's."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_here = _os.path.abspath(__file__)
for _p in _sys.path:
    _cand = _os.path.join(_p, "models", "functions", "funcs.py")
    if _os.path.isfile(_cand) and _os.path.abspath(_cand) != _here:
        _spec = _ilu.spec_from_file_location("_reference_models_functions_funcs", _cand)
        _mod = _ilu.module_from_spec(_spec)
        _spec.loader.exec_module(_mod)
        globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("_")})
        break
from planerecnet_b200.models.functions.funcs import bias_init_with_prob  # noqa: E402,F401

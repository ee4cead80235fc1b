"""Drop-in ``quant`` package: quantizer / int_linear / int_matmul are served from here.

Put ``llama3-quantization_b200/`` in front of the reference checkout on ``sys.path`` and the
reference's own ``models/int_llama_layer.py``, ``quant/omniquant.py``, ``main.py`` ... import
these three modules instead of theirs.  The reference's remaining ``quant/*`` modules
(``omni_norm``, ``omniquant``, ``utils``) are out of scope for the rewrite and are reached
through ``__path__``: set ``Q100_REF_PATH`` (or have the reference checkout on ``sys.path``)
and they resolve to the reference's files.
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))


def _reference_quant_dirs():
    cands = []
    env = _os.environ.get("Q100_REF_PATH")
    if env:
        cands.append(_os.path.join(env, "quant"))
    for p in _sys.path:
        if p and _os.path.abspath(p) != _os.path.dirname(_here):
            d = _os.path.join(p, "quant")
            if _os.path.isfile(_os.path.join(d, "omniquant.py")):
                cands.append(d)
    out = []
    for d in cands:
        d = _os.path.abspath(d)
        if d != _here and _os.path.isdir(d) and d not in out:
            out.append(d)
    return out


for _d in _reference_quant_dirs():
    if _d not in __path__:
        __path__.append(_d)

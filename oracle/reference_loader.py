"""Import the reference's own model code, unchanged, from /root/reference.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has
no /root/reference); used by oracle/make_golden.py to generate tests/golden/
and by the ``not gpu`` tests that pin oracle/dost_oracle.py against the live
reference when it is present.
"""
from __future__ import annotations

import importlib
import os
import sys

from . import shims

REFERENCE_ROOT = os.environ.get("DOST_REFERENCE_ROOT", "/root/reference")
_CACHE = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "embedder_eDOS", "DOSTransformer.py"))


def load():
    """Returns (DOSTransformer, DOSTransformer_phonon, TransformerEncoder) classes of the reference."""
    if "classes" in _CACHE:
        return _CACHE["classes"]
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    shims.install()
    names = ("layers", "embedder_eDOS", "embedder_phDOS")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in names}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        edos = importlib.import_module("embedder_eDOS.DOSTransformer")
        ph = importlib.import_module("embedder_phDOS.DOSTransformer_phonon")
        lay = importlib.import_module("layers")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k.split(".")[0] in names]:
            del sys.modules[k]
        sys.modules.update(saved)
    _CACHE["classes"] = (edos.DOSTransformer, ph.DOSTransformer_phonon, lay.TransformerEncoder)
    return _CACHE["classes"]

"""Import the reference's own model code, unchanged, from /root/reference or from its staged copy oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  /root/reference exists only in the build container; oracle/build_ref.py stages the files of
the hot path under oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot), byte-for-byte.  Used by
oracle/make_golden.py to generate tests/golden/, by the ``not gpu`` tests that pin oracle/dost_oracle.py against the
live reference, and by bench.py's CPU legs (``--impl reference`` / ``cpu_baseline``, kind "reference").
"""
from __future__ import annotations

import importlib
import os
import sys

from . import shims

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root() -> str:
    env = os.environ.get("DOST_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(root, "embedder_eDOS", "DOSTransformer.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()
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

"""Recipe that stages the reference's own model files under oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is staging the files of the hot path where the
GPU box can import them: `/root/reference` exists only in the build container, `oracle/_ref/` travels with the snapshot
like the built .so.  Nothing is copied into the git history (oracle/_ref/ is in .gitignore); the files are byte-for-byte
the reference's (a SHA-256 manifest is written next to them) and are imported UNCHANGED behind oracle/shims.py, which
stands in for the three un-installed third-party packages (torch_scatter, torch_geometric.utils.to_dense_batch, e3nn).

    python -m oracle.build_ref            # called by __graft_entry__.build() when /root/reference is present

Used by: `bench.py --impl reference` and `cpu_baseline` (kind "reference": the reference's own nn.Modules timed on the
box's host cores), and the `not gpu` tests that pin oracle/dost_oracle.py against the live reference.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DOST_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "embedder_eDOS/__init__.py", "embedder_eDOS/DOSTransformer.py",
    "embedder_phDOS/__init__.py", "embedder_phDOS/DOSTransformer_phonon.py",
    "layers/__init__.py", "layers/transformer.py", "layers/multihead_attention.py",
]


def build(verbose: bool = False) -> str | None:
    """Stages the files; returns the destination, or None when the reference is not present (GPU box)."""
    if not os.path.isfile(os.path.join(SRC, FILES[1])):
        return DST if os.path.isfile(os.path.join(DST, FILES[1])) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files under {DST}")
    return DST


if __name__ == "__main__":
    print(build(verbose=True))

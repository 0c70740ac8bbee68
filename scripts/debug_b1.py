"""B=1 forward at hidden=128 on the tensor-core path: where do non-finite values appear, under which switch?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200.collate import PackedCrystals, split_batch
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch

dev = "cuda"
src = make_edos_batch(21, seed=41)
store = PackedCrystals.from_graphs(split_batch(src), device=dev)
torch.manual_seed(1)
model = DOSTransformer(2, 1, 200, 41, 2, 128, torch.device(dev), 0.0).to(dev).eval()
big = int(store.node_count.argmax())
print("largest crystal", big, "nodes", store.node_count[big])
for sw in ["", "DOST_NO_XATTN_TC", "DOST_NO_ATTNPLANES", "DOST_NO_HEADSPLIT", "DOST_NO_LINPLANES", "DOST_NO_FFNBLOCK",
           "DOST_NO_EDGEBLOCK", "DOST_NO_LNVEC"]:
    for k in list(os.environ):
        if k.startswith("DOST_NO_"):
            del os.environ[k]
    if sw:
        os.environ[sw] = "1"
    L.reload_switches()
    for pc in (False, True):
        model.per_crystal_eval = pc
        for ids in ([big], [0], [big, 0], [3, 5, big]):
            with torch.no_grad():
                dg, x, ds = model(store.collate(ids))
            print(f"switch={sw or '-':20s} per_crystal={pc!s:5s} ids={ids!s:12s} finite: dg={bool(torch.isfinite(dg).all())} "
                  f"x={bool(torch.isfinite(x).all())} ds={bool(torch.isfinite(ds).all())}", flush=True)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dostransformer_b200 import _lib, ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch
DEV = "cuda"
torch.manual_seed(7)
m = DOSTransformer(2, 1, 200, 41, 2, 128, torch.device(DEV), 0.0).to(DEV)
g = make_edos_batch(4, seed=55, sizes=torch.tensor([150, 301, 37, 222]), K=24).to(DEV)
def run(prec):
    m.precision = prec
    m.train(); m.zero_grad(set_to_none=True)
    dg, x, ds = m(g)
    loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos"); loss.backward()
    return {k: p.grad.double().clone() for k, p in m.named_parameters() if p.grad is not None}
ref = run("fp32")
toggles = ["", "DOST_NO_FFNBLOCK", "DOST_NO_EDGEBLOCK", "DOST_NO_ATTNPLANES", "DOST_NO_HEADSPLIT", "DOST_NO_LINPLANES+DOST_NO_HEADSPLIT",
           "DOST_NO_FFNBLOCK+DOST_NO_EDGEBLOCK+DOST_NO_ATTNPLANES+DOST_NO_HEADSPLIT+DOST_NO_LINPLANES"]
for t in toggles:
    for k in list(os.environ):
        if k.startswith("DOST_NO_"): del os.environ[k]
    for k in t.split("+"):
        if k: os.environ[k] = "1"
    _lib.reload_switches()
    got = run("bf16x3")
    errs = sorted(((((got[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-30)).item(), k) for k in ref), reverse=True)
    print(f"[{t or 'all on'}] worst:", ", ".join(f"{k}={e:.1e}" for e, k in errs[:4]))

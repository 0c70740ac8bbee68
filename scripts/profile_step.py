"""One training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).  B=<crystals>, WORKLOAD=edos|phonon."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.embedder_phDOS.DOSTransformer_phonon import DOSTransformer_phonon
from dostransformer_b200.synthetic import make_edos_batch, make_phonon_batch

B = int(os.environ.get("B", "512"))
dev = torch.device("cuda")
torch.manual_seed(0)
MODE = os.environ.get("WORKLOAD", "edos")
if MODE == "phonon":      # main_phDOS.py defaults: float64
    model = DOSTransformer_phonon(3, 2, 118, 4, 256, dev, 0.0).to(dev).double().train()
    g = make_phonon_batch(B, seed=1000).to(dev)
else:
    model = DOSTransformer(3, 2, 200, 41, 2, 256, dev, 0.0, precision=os.environ.get("DOST_PRECISION", "bf16x3")).to(dev).train()
    g = make_edos_batch(B, seed=2000).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    dg, _, ds = model(g)
    loss = ops.dos_loss(dg, ds, g.y_ft if MODE == "edos" else g.phdos, mode=MODE, beta=1.0)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")

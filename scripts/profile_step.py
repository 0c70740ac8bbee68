"""One eDOS training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch

B = int(os.environ.get("B", "512"))
dev = torch.device("cuda")
torch.manual_seed(0)
model = DOSTransformer(3, 2, 200, 41, 2, 256, dev, 0.0, precision=os.environ.get("DOST_PRECISION", "bf16x3")).to(dev).train()
g = make_edos_batch(B, seed=2000).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    dg, _, ds = model(g)
    loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos", beta=1.0)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")

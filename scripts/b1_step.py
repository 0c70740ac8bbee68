"""One fwd+bwd of a single-crystal batch on the tensor-core path (run under compute-sanitizer: the ragged attention GEMMs
of a one-crystal batch must neither read nor write past their buffers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch

dev = "cuda"
torch.manual_seed(0)
model = DOSTransformer(1, 1, 200, 41, 2, 128, torch.device(dev), 0.0).to(dev).train()
for seed in (3, 4):
    g = make_edos_batch(1, seed=seed).to(dev)
    dg, x, ds = model(g)
    loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos", beta=1.0)
    loss.backward()
    torch.cuda.synchronize()
    print("loss", float(loss), "finite grads", all(bool(torch.isfinite(p.grad).all()) for p in model.parameters() if p.grad is not None))

"""Finds reference cycles left behind by one eager training step (they keep activations alive until the cyclic GC runs)."""
import gc
import os
import sys
from collections import Counter

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch

dev = torch.device("cuda")
torch.manual_seed(0)
model = DOSTransformer(3, 2, 200, 41, 2, 256, dev, 0.0).to(dev).train()
g = make_edos_batch(64, seed=2000).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    dg, _, ds = model(g)
    loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos", beta=1.0)
    loss.backward()
    return float(loss)


for _ in range(2):
    step()
gc.collect()
gc.disable()
torch.cuda.synchronize()
m0 = torch.cuda.memory_allocated()
step()
torch.cuda.synchronize()
m1 = torch.cuda.memory_allocated()
print("allocated after one step (MB), before/after:", m0 / 1e6, m1 / 1e6)
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
print("unreachable objects:", n)
types = Counter(type(o).__name__ for o in gc.garbage)
print(types.most_common(25))
big = [(o.numel() * o.element_size() / 1e6, tuple(o.shape), type(o.grad_fn).__name__ if o.grad_fn is not None else None)
       for o in gc.garbage if torch.is_tensor(o)]
big.sort(reverse=True)
print("tensors in cycles:", len(big), "total MB", sum(b[0] for b in big))
for b in big[:25]:
    print("  ", b)
for o in gc.garbage:
    if type(o).__name__ in ("LinearSpec", "CrystalGraph", "Planes", "_Box", "CSR", "RowMap"):
        print("obj", type(o).__name__, [type(r).__name__ for r in gc.get_referrers(o)][:6])

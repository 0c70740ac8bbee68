"""cProfile of the host side of a training step at a CPU-bound batch size (B=8)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dostransformer_b200 import ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch
dev = torch.device("cuda")
torch.manual_seed(0)
model = DOSTransformer(3, 2, 200, 41, 2, 256, dev, 0.0).to(dev).train()
g = make_edos_batch(int(os.environ.get("B", "8")), seed=2000).to(dev)
def step():
    model.zero_grad(set_to_none=True)
    dg, _, ds = model(g)
    loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos", beta=1.0)
    loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)

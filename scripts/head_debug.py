import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dostransformer_b200 import ops, nn_core
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch
DEV = "cuda"
torch.manual_seed(0)
m = DOSTransformer(3, 2, 200, 41, 2, 256, torch.device(DEV), 0.0).to(DEV)
g = make_edos_batch(24, seed=77).to(DEV)
logs = {}
orig_linear, orig_ln, orig_self, orig_cross = ops.linear, ops.layer_norm, ops.self_attention, ops.cross_attention
def wrap(name, fn):
    def f(*a, **k):
        out = fn(*a, **k)
        o = out[0] if isinstance(out, tuple) else out
        logs[cfg].append((name, tuple(o.shape), o.detach().double().norm().item()))
        return out
    return f
ops.linear = wrap("linear", orig_linear); ops.layer_norm = wrap("ln", orig_ln)
ops.self_attention = wrap("self", orig_self); ops.cross_attention = wrap("cross", orig_cross)
for cfg, env in [("A", {"DOST_NO_HEADSPLIT": "1"}), ("B", {}), ("B2", {})]:
    os.environ.pop("DOST_NO_HEADSPLIT", None)
    os.environ.update(env)
    logs[cfg] = []
    m.precision = "bf16x3"
    m.train()
    dg, x, ds = m(g)
    loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos")
    loss.backward()
    torch.cuda.synchronize()
    logs[cfg].append(("dg", tuple(dg.shape), dg.double().norm().item()))
ia = ib = 0
A, B = logs["B2"], logs["B"]
print(len(A), len(B))
for i in range(max(len(A), len(B))):
    a = A[i] if i < len(A) else None
    b = B[i] if i < len(B) else None
    print(i, a, b)

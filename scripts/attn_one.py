"""Runs the fused attention forward at the step's shapes a few times, for ncu captures.
usage: attn_one.py <self|cross> [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import ops
from dostransformer_b200.synthetic import make_edos_batch

kind = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, T, H = 512, 201, 256
S = 2 * B
dev = "cuda"
with ops.precision("bf16x3"):
    q, k, r = (torch.randn(S, T, H, device=dev) for _ in range(3))
    q._dost_planes = ops.split_planes(q.view(S * T, H))
    k._dost_planes = ops.split_planes(k.view(S * T, H))
    qq = q.clone().requires_grad_(True)
    qq._dost_planes = q._dost_planes
    if kind == "self":
        fn = lambda: ops.self_attention(qq, k, r)        # training forward: the probability planes are saved
    else:
        g = make_edos_batch(B, seed=2000, T=T)
        gr = ops.build_graph(g.edge_index.to(dev), g.batch.to(dev), g.system.to(dev), nmax_hint=g.max_num_nodes)
        kv, ph = torch.randn(gr.N, H, device=dev), torch.randn(H, device=dev)
        fn = lambda: ops.cross_attention(qq, kv, ph, r, gr, S)
    for _ in range(iters):
        fn()
torch.cuda.synchronize()
print("ok", kind)

"""Times the attention forward (and forward+backward) of the step's shapes: fused single kernel vs GEMM + softmax + GEMM.
usage: attn_time.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops
from dostransformer_b200.synthetic import make_edos_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T, H = 201, 256
dev = "cuda"


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (L.launch_count() - n0) / iters


g = make_edos_batch(B, seed=2000, T=T)
gr = ops.build_graph(g.edge_index.to(dev), g.batch.to(dev), g.system.to(dev), nmax_hint=g.max_num_nodes)
print("B", B, "N", gr.N, "nmax", g.max_num_nodes)
with ops.precision("bf16x3"):
    for S in (B, 2 * B):
        q = torch.randn(S, T, H, device=dev)
        k = torch.randn(S, T, H, device=dev)
        r = torch.randn(S, T, H, device=dev)
        q._dost_planes = ops.split_planes(q.view(S * T, H))
        k._dost_planes = ops.split_planes(k.view(S * T, H))
        kv = torch.randn(gr.N, H, device=dev)
        ph = torch.randn(H, device=dev)
        for sw in ("", "1"):
            if sw:
                os.environ["DOST_NO_ATTN_FUSED"] = "1"
            else:
                os.environ.pop("DOST_NO_ATTN_FUSED", None)
            L.reload_switches()
            tag = "unfused" if sw else "fused  "
            with torch.no_grad():
                ms, nl = timeit(lambda: ops.self_attention(q, k, r))
                print(f"S={S} self  fwd(no grad) {tag} {ms:.3f} ms  {nl:.0f} launches")
                ms, nl = timeit(lambda: ops.cross_attention(q, kv, ph, r, gr, S))
                print(f"S={S} cross fwd(no grad) {tag} {ms:.3f} ms  {nl:.0f} launches")
            qq, kk, kvv = q.clone().requires_grad_(True), k.clone().requires_grad_(True), kv.clone().requires_grad_(True)
            qq._dost_planes, kk._dost_planes = q._dost_planes, k._dost_planes

            def fb_self():
                o = ops.self_attention(qq, kk, r)
                o.backward(r)
                qq.grad = kk.grad = None

            def fb_cross():
                o = ops.cross_attention(qq, kvv, ph, r, gr, S)
                o.backward(r)
                qq.grad = kvv.grad = None
            ms, nl = timeit(fb_self)
            print(f"S={S} self  fwd+bwd      {tag} {ms:.3f} ms  {nl:.0f} launches")
            ms, nl = timeit(fb_cross)
            print(f"S={S} cross fwd+bwd      {tag} {ms:.3f} ms  {nl:.0f} launches")

import sys, os
sys.path.insert(0, os.getcwd())
import torch
from dostransformer_b200 import ops
from dostransformer_b200.synthetic import make_edos_batch
dev="cuda"
g = make_edos_batch(512, seed=2000, T=201)
gr = ops.build_graph(g.edge_index.to(dev), g.batch.to(dev), g.system.to(dev))
for W in (256, 512):
    dst = torch.empty(gr.N, W, device=dev)
    big = [torch.randn(gr.E, W, device=dev) for _ in range(max(1, int(300e6 // (gr.E * W * 4))))]
    for rep in range(2):
        for i in range(3): ops.segment_reduce_raw(big[i % len(big)], gr.by_dst.rowptr, gr.by_dst.perm, gr.N, out=dst)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 30
        for i in range(n): ops.segment_reduce_raw(big[i % len(big)], gr.by_dst.rowptr, gr.by_dst.perm, gr.N, out=dst)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        nbytes = 4.0 * W * (gr.E + gr.N) + 4.0 * gr.E + 4.0 * (gr.N + 1)
        print(f"W={W} {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s = {nbytes/ms/1e6/6459:.2f}")

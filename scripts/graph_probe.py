"""Eager vs CUDA-graph replay step time of the eDOS training step at several per-GPU batch sizes (one GPU)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dostransformer_b200 import ops  # noqa: E402
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer  # noqa: E402
from dostransformer_b200.graphed import GraphedStep  # noqa: E402
from dostransformer_b200.synthetic import make_edos_batch, pad_edos_batch  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
model = DOSTransformer(3, 2, 200, 41, 2, 256, dev, 0.0).to(dev).train()
out = {}
for B in [int(b) for b in (sys.argv[1:] or ["8", "64", "512"])]:
    raw = [make_edos_batch(B, seed=2000 + i) for i in range(3)]
    padded = [pad_edos_batch(g).to(dev) for g in raw]
    gs = [g.to(dev) for g in raw]

    def eager(g):
        model.zero_grad(set_to_none=True)
        dg, _, ds = model(g)
        loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos", beta=1.0)
        loss.backward()
        return loss

    def timeit(fn, batches, n=12):
        for g in batches:
            fn(g)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(n):
            fn(batches[i % len(batches)])
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n, (time.perf_counter() - t0) / n * 1e3

    e_dev, e_wall = timeit(eager, gs)
    step = GraphedStep(model, "edos")
    g_dev, g_wall = timeit(step, gs)
    stepp = GraphedStep(model, "edos")
    p_dev, p_wall = timeit(stepp, padded)
    out[B] = {"eager_ms": e_dev, "eager_wall_ms": e_wall, "graph_ms": g_dev, "graph_wall_ms": g_wall,
              "graph_padded_ms": p_dev, "captures": step.captures, "captures_padded": stepp.captures,
              "launches_per_step": step.launches / max(step.replays, 1),
              "crystals_per_s_eager": B / e_dev * 1e3, "crystals_per_s_graph": B / g_dev * 1e3}
    print(B, json.dumps(out[B]), flush=True)

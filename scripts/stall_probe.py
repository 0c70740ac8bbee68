"""Looks for sporadic stalls of the eager step: per-step host wall time and device time, cudaMalloc counts."""
import gc
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.optim import AdamW
from dostransformer_b200.synthetic import make_edos_batch, CrystalBatch

dev = torch.device("cuda")
torch.manual_seed(0)
model = DOSTransformer(3, 2, 200, 41, 2, 256, dev, 0.0).to(dev).train()
host = [make_edos_batch(512, seed=2000 + i).pin_memory() for i in range(3)]
model.max_num_nodes = max(h.max_num_nodes for h in host)
resident = [h.clone().to(dev) for h in host]
opt = AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)


def to_dev(g):
    return CrystalBatch(**{k: (getattr(g, k).to(dev, non_blocking=True) if torch.is_tensor(getattr(g, k)) else getattr(g, k)) for k in g.keys()})


def step(g):
    model.zero_grad(set_to_none=True)
    dg, _, ds = model(g)
    loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos", beta=1.0)
    loss.backward()
    return loss.detach()


def run(name, fn, n=24):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    gc.collect(); gc.freeze(); gc.disable()
    s0 = torch.cuda.memory_stats()
    host_ms, ev = [], [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        t = time.perf_counter()
        fn(i)
        host_ms.append((time.perf_counter() - t) * 1e3)
        ev[i + 1].record()
    torch.cuda.synchronize()
    gc.enable()
    s1 = torch.cuda.memory_stats()
    dev_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print(name, "host", [round(x, 1) for x in host_ms])
    print(name, "dev ", [round(x, 1) for x in dev_ms])
    print(name, "cudaMalloc calls", s1["num_device_alloc"] - s0["num_device_alloc"], "frees", s1["num_device_free"] - s0["num_device_free"],
          "retries", s1["num_alloc_retries"] - s0["num_alloc_retries"], "reserved GB", s1["reserved_bytes.all.current"] / 1e9, flush=True)


run("resident", lambda i: step(resident[i % 3]))
run("e2e", lambda i: step(to_dev(host[i % 3])).item())
run("opt", lambda i: (step(resident[i % 3]), opt.step()))
os.environ["DOST_GEMM_TMA_EPI"] = "0"
run("e2e_noTMA", lambda i: step(to_dev(host[i % 3])).item())
run("opt_noTMA", lambda i: (step(resident[i % 3]), opt.step()))

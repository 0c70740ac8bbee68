"""CPU timing of the reference's train step on synthetic batches (bench.py's `--impl reference` and `cpu_baseline`).

TEST INFRASTRUCTURE ONLY: imported by bench.py's CPU legs, never by the product package.

kind "reference": the reference's OWN nn.Modules (embedder_eDOS/DOSTransformer.py, embedder_phDOS/
DOSTransformer_phonon.py, layers/) imported unchanged from /root/reference or its staged copy oracle/_ref/ behind
oracle/shims.py, driven by the launchers' step lines (main_eDOS.py:106-126, main_phDOS.py:101-116: forward, loss,
backward; no optimizer, as the metric is defined).  kind "port": oracle/dost_oracle.py, the functional restatement pinned
against the reference by tests/golden, when neither copy of the reference is present.
"""
from __future__ import annotations

import os
import statistics
import time

import torch


def _edos_loss(dg, ds, g, beta=1.0):          # main_eDOS.py:111-123
    zero = torch.tensor(0, dtype=g.y_ft.dtype)
    y_ft = torch.where(g.y_ft < 0, zero, g.y_ft)
    y = y_ft.reshape(len(g.mp_id), -1)
    return torch.sqrt(((y - dg) ** 2).mean(dim=1)).mean() + beta * torch.sqrt(((y - ds) ** 2).mean(dim=1)).mean()


def _phonon_loss(dg, ds, g, beta=1.0):        # main_phDOS.py:109-114
    crit = torch.nn.MSELoss()
    return torch.sqrt(crit(dg, g.phdos)).mean() + beta * torch.sqrt(crit(ds, g.phdos)).mean()


def make_stepper(workload: str, hidden: int = 256, layers: int = 3, t_layers: int = 2, T: int = None):
    """Returns (step(g) -> loss, kind).  workload: "edos" (fp32) | "phonon" (fp64, main_phDOS.py:15-16)."""
    from . import reference_loader
    cpu = torch.device("cpu")
    dtype = torch.float32 if workload == "edos" else torch.float64
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        torch.manual_seed(0)
        if reference_loader.available():
            EDOS, PHONON, _ = reference_loader.load()
            if workload == "edos":
                model, loss_fn = EDOS(layers, t_layers, 200, 41, 2, hidden, cpu, 0.0), _edos_loss
            else:
                model, loss_fn = PHONON(layers, t_layers, 118, 4, hidden, cpu, 0.0), _phonon_loss
            if T is not None and T != model.embeddings.weight.shape[0]:
                raise ValueError("the reference hard-codes its energy-grid length")
            model.train()

            def step(g):
                model.zero_grad(set_to_none=True)
                dg, _, ds = model(g)
                loss = loss_fn(dg, ds, g)
                loss.backward()
                return loss

            return step, "reference"
        from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
        from dostransformer_b200.embedder_phDOS.DOSTransformer_phonon import DOSTransformer_phonon
        from . import dost_oracle as O
        if workload == "edos":
            sd = O.state_dict_of(DOSTransformer(layers, t_layers, 200, 41, 2, hidden, "cpu", 0.0))
            fwd, lossf, tkey = O.edos_forward, O.edos_loss, "y_ft"
        else:
            sd = O.state_dict_of(DOSTransformer_phonon(layers, t_layers, 118, 4, hidden, "cpu", 0.0).double())
            fwd, lossf, tkey = O.phonon_forward, O.phonon_loss, "phdos"

        def step(g):
            return O.run_train_step(fwd, lossf, sd, g, getattr(g, tkey))[1]

        return step, "port"
    finally:
        torch.set_default_dtype(prev)


def throughput(workload: str, batch, steps: int, warmup: int, threads: int):
    """fwd+bwd on `batch` (CPU tensors); returns dict(value crystals/s, ms median, total seconds, kind, cores)."""
    torch.set_num_threads(int(threads))
    step, kind = make_stepper(workload)
    B = int(batch.system.numel())
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(batch)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return {"value": B / med, "ms_per_step": med * 1e3, "seconds": sum(times), "kind": kind, "cores": int(threads),
            "host_cores": os.cpu_count() or 1, "steps": steps, "crystals_per_step": B}

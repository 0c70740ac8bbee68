"""Device-side evaluation / inference sweep (BASELINE config 5).

``evaluate`` mirrors ``utils.test`` (utils.py:61-112) and ``utils.test_phonon`` (utils.py:117-143): forward in eval mode
under ``no_grad``, targets (and, for eDOS, system predictions) clamped at 0, per-crystal MSE / RMSE / MAE / R^2 averaged
over the data set, the pooled atom embeddings ``scatter_sum(x, batch)`` (utils.py:91) and the predictions returned for
saving.  The reference's loaders use ``batch_size=1``: every crystal is evaluated without padding.  Here crystals are
evaluated many per launch with ``model.per_crystal_eval = True`` (no phantom keys, identical per-crystal results) and
nothing is read back to the host until the end of the sweep.
"""
from __future__ import annotations

from typing import Iterable

import torch

from . import ops


@torch.no_grad()
def evaluate(model, batches: Iterable, mode: str = "edos", device=None):
    """Returns (rmse, mse, mae, r2, [mp_id, preds, y, embeddings]) like utils.test; all four metrics are python floats."""
    was_training, was_pc = model.training, model.per_crystal_eval
    model.eval()
    model.per_crystal_eval = True
    per_all, preds_all, y_all, emb_all, ids = [], [], [], [], []
    try:
        for g in batches:
            if device is not None:
                g = g.to(device)
            dos_global, x, dos_system = model(g)
            B, T = dos_system.shape
            if mode == "edos":
                target, pred, clamp = g.y_ft.reshape(B, T), dos_system, True
            else:
                target, pred, clamp = g.phdos.reshape(B, T), dos_system, False
            per, _ = ops.eval_metrics(pred, target, clamp_pred=clamp)
            graph_ptr = ops.csr_build(ops.to_i32(g.batch), B)[0]
            graph_ptr.perm = None
            emb_all.append(ops.segment_reduce_raw(x.contiguous(), graph_ptr.rowptr, None, B))
            per_all.append(per)
            preds_all.append(pred.clamp_min(0) if clamp else pred)
            y_all.append(target.clamp_min(0) if clamp else target)
            ids += list(getattr(g, "mp_id", []))
    finally:
        model.per_crystal_eval = was_pc
        model.train(was_training)
    per = torch.cat(per_all)                       # [n_crystals, 4] = mse, rmse, mae, r2
    mean = per.double().mean(0).cpu()              # the only device->host read of the sweep (besides the returned arrays)
    preds_y = [ids, torch.cat(preds_all).cpu().numpy(), torch.cat(y_all).cpu().numpy(), torch.cat(emb_all).cpu().numpy()]
    return float(mean[1]), float(mean[0]), float(mean[2]), float(mean[3]), [preds_y]


def sweep_plan(node_count, batch_size: int, world: int = 1):
    """Host-side plan of an inference sweep over a packed store (BASELINE config 5: "sharded contiguous-by-length"):
    crystals sorted by atom count (stable), cut into batches of ``batch_size`` neighbours in that order (a batch's
    padding-free work is then uniform), batches dealt round-robin to the ranks.  Returns per rank a list of int64 id
    tensors; every crystal appears exactly once."""
    import numpy as np
    order = np.argsort(np.asarray(node_count), kind="stable")
    batches = [torch.from_numpy(order[i:i + batch_size].astype(np.int64)) for i in range(0, len(order), batch_size)]
    return [batches[r::world] for r in range(world)]


@torch.no_grad()
def sweep(model, store, batch_size: int = 512, rank: int = 0, world: int = 1, clamp: bool = True):
    """Forward-only DOS prediction of every crystal of ``store`` (a collate.PackedCrystals) assigned to this rank: batches
    are assembled on the device from crystal ids, evaluated per crystal (``per_crystal_eval``: no padding coupling between
    crystals, like the reference's batch_size-1 test loaders, utils.py:61-112) and the predictions are scattered into one
    device buffer; nothing is read back per batch.  Returns (ids [n_local] int64 on the device, dos_system [n_local, T],
    dos_global [n_local, T]), rows in the order of ``ids``; predictions clamped at 0 like utils.py:76 if ``clamp``."""
    plan = sweep_plan(store.node_count, batch_size, world)[rank]
    was_training, was_pc = model.training, model.per_crystal_eval
    model.eval()
    model.per_crystal_eval = True
    ids_all, sys_all, glob_all = [], [], []
    try:
        for ids in plan:
            g = store.collate(ids)
            dos_global, _, dos_system = model(g)
            ids_all.append(ids)
            sys_all.append(dos_system.clamp_min(0) if clamp else dos_system)
            glob_all.append(dos_global.clamp_min(0) if clamp else dos_global)
    finally:
        model.per_crystal_eval = was_pc
        model.train(was_training)
    if not ids_all:
        dev = store.device
        return torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros(0, 0, device=dev), torch.zeros(0, 0, device=dev)
    return torch.cat(ids_all).to(store.device), torch.cat(sys_all), torch.cat(glob_all)

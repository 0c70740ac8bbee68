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
            y_all.append(target.clamp_min(0))
            ids += list(getattr(g, "mp_id", []))
    finally:
        model.per_crystal_eval = was_pc
        model.train(was_training)
    per = torch.cat(per_all)                       # [n_crystals, 4] = mse, rmse, mae, r2
    mean = per.double().mean(0).cpu()              # the only device->host read of the sweep (besides the returned arrays)
    preds_y = [ids, torch.cat(preds_all).cpu().numpy(), torch.cat(y_all).cpu().numpy(), torch.cat(emb_all).cpu().numpy()]
    return float(mean[1]), float(mean[0]), float(mean[2]), float(mean[3]), [preds_y]

"""CPU restatement of the DOSTransformer / DOSTransformer_phonon hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the parity checker for the
CUDA path and the ``cpu_baseline`` of bench.py.  Plain PyTorch on the CPU,
functional (weights come in as a ``state_dict``-style mapping with the
reference's parameter names), written in the padded/dense form the reference
uses so that it restates the reference's arithmetic, not the CUDA design.

Pinned by tests/test_oracle.py against tests/golden/*.pt, which hold outputs of
the reference's own model code (oracle/make_golden.py) -- and against the live
reference when /root/reference is present.  The third-party scatter /
to_dense_batch / e3nn primitives are restated from their published semantics
(oracle/shims.py): "parity unpinned" for those.

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Mapping[str, torch.Tensor]


# ------------------------------------------------------------------ third-party primitives
def segment_sum(src: torch.Tensor, index: torch.Tensor, size: int) -> torch.Tensor:
    """torch_scatter.scatter_sum(dim=0) -- call sites DOSTransformer.py:158,187."""
    return src.new_zeros((size,) + tuple(src.shape[1:])).index_add_(0, index, src)


def segment_mean(src: torch.Tensor, index: torch.Tensor, size: int) -> torch.Tensor:
    """torch_scatter.scatter_mean(dim=0) -- DOSTransformer_phonon.py:209; empty rows give 0."""
    cnt = torch.bincount(index, minlength=size).clamp(min=1).to(src.dtype)
    return segment_sum(src, index, size) / cnt[:, None]


def pad_crystals(x: torch.Tensor, batch: torch.Tensor, max_num_nodes: Optional[int] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """torch_geometric.utils.to_dense_batch -- DOSTransformer.py:61.  Returns
    ([B, Nmax, H] zero padded, atoms per crystal, Nmax).  ``max_num_nodes`` is PyG's argument of the same name
    (the reference never passes it; the data-parallel tests use it to impose the global padding length)."""
    B = int(batch.max()) + 1
    n = torch.bincount(batch, minlength=B)
    nmax = int(n.max()) if max_num_nodes is None else int(max_num_nodes)
    start = torch.cumsum(n, 0) - n
    slot = torch.arange(batch.numel(), device=batch.device) - start[batch] + batch * nmax
    dense = x.new_zeros(B * nmax, x.shape[1])
    dense[slot] = x
    return dense.view(B, nmax, x.shape[1]), n, nmax


def phonon_edge_features(edge_vec: torch.Tensor) -> torch.Tensor:
    """DOSTransformer_phonon.py:75-77: smooth_cutoff(|v|/4) * [1, sqrt(3) v/|v|]."""
    length = edge_vec.norm(dim=1)
    unit = edge_vec / length.clamp_min(1e-12)[:, None]
    sh = torch.cat([torch.ones_like(length)[:, None], math.sqrt(3.0) * unit], dim=1)
    u = 2.0 * (length / 4.0 - 1.0)
    cut = (1.0 - torch.cos(math.pi * u)) / 2.0
    cut = torch.where(u > 0, torch.zeros_like(cut), cut)
    cut = torch.where(u < -1, torch.ones_like(cut), cut)
    return cut[:, None] * sh


# ------------------------------------------------------------------ small building blocks
def _lin(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _ln(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    w = p[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, p[name + ".bias"], 1e-5)


def _mlp_prelu(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    """Sequential(Linear, PReLU, Linear) -- DOSTransformer.py:103-105."""
    return _lin(p, name + ".2", F.prelu(_lin(p, name + ".0", x), p[name + ".1.weight"]))


def _mlp_ln_prelu(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    """Sequential(Linear, LayerNorm, PReLU, Linear) -- DOSTransformer.py:171,182."""
    h = _ln(p, name + ".1", _lin(p, name + ".0", x))
    return _lin(p, name + ".3", F.prelu(h, p[name + ".2.weight"]))


def attention(q: torch.Tensor, kv: torch.Tensor, drop_p: float = 0.0, training: bool = False) -> torch.Tensor:
    """layers/multihead_attention.py:49-76 with num_heads=1: no projections,
    scaling = embed_dim**-0.5, softmax in fp32 regardless of the input dtype,
    key == value.  q [S, Lq, H], kv [S, Lk, H] (batch-first here)."""
    scores = torch.bmm(q, kv.transpose(1, 2)) * (q.shape[-1] ** -0.5)
    prob = F.softmax(scores.float(), dim=-1).type_as(scores)
    prob = F.dropout(prob, p=drop_p, training=training)
    return torch.bmm(prob, kv)


def encoder_stack(p: Params, name: str, x: torch.Tensor, kv: torch.Tensor, n_layers: int,
                  drop_p: float = 0.0, training: bool = False) -> torch.Tensor:
    """layers/transformer.py:46-79 (stack) and :120-150 (layer), pre-LN form.
    Keys/values are the ORIGINAL ``kv`` in every layer, normalised by that
    layer's layer_norms[0]; only the query stream evolves."""
    for i in range(n_layers):
        pre = f"{name}.layers.{i}"
        a = attention(_ln(p, pre + ".layer_norms.0", x), _ln(p, pre + ".layer_norms.0", kv), drop_p, training)
        x = x + a
        h = F.relu(_lin(p, pre + ".fc1", _ln(p, pre + ".layer_norms.1", x)))
        x = x + _lin(p, pre + ".fc2", h)
    return _ln(p, name + ".layer_norm", x)


def _message_passing(p: Params, x, e, row, col, n_layers: int, mean: bool):
    """Processor loop -- DOSTransformer.py:56-59,137-148,173-175,184-190."""
    N = x.shape[0]
    for i in range(n_layers):
        pre = f"stacked_processor.{i}"
        e_out = _mlp_ln_prelu(p, pre + ".edge_model.edge_mlp", torch.cat([x[row], x[col], e], dim=1))
        agg = segment_mean(e_out, col, N) if mean else segment_sum(e_out, col, N)
        x_out = _mlp_ln_prelu(p, pre + ".node_model.node_mlp_2", torch.cat([x, agg], dim=1))
        x = x + x_out
        e = e + e_out
    return x, e


def _dos_heads(p: Params, x, batch, system, energies_tok, graph, prompt_name, t_layers, drop_p, training,
               max_num_nodes=None):
    """Everything after the GNN -- DOSTransformer.py:61-91 / DOSTransformer_phonon.py:86-117."""
    T = energies_tok.shape[0]
    dense, n, nmax = pad_crystals(x, batch, max_num_nodes)      # [B, Nmax, H]; padded rows are zero
    B = dense.shape[0]
    q0 = energies_tok[None].expand(B, T, -1)
    energies = encoder_stack(p, "transformer", q0, dense, t_layers, drop_p, training)
    g = graph[:, None, :].expand(B, T, -1)
    prompt = p[prompt_name + ".weight"][system][:, None, :].expand(B, T, -1)

    def branch(inp):
        h = encoder_stack(p, "transformer_self", inp, inp, t_layers, drop_p, training)
        h = encoder_stack(p, "transformer_source", h, dense, t_layers, drop_p, training)
        return _lin(p, "out_layer", h).squeeze(2)               # [B, T]

    dos_global = branch(F.leaky_relu(_lin(p, "fc", torch.cat([energies, g], dim=2))))
    dos_system = branch(F.leaky_relu(_lin(p, "fc_prompt", torch.cat([energies, g, prompt], dim=2))))
    return dos_global, dos_system


def edos_forward(p: Params, g, *, training: bool = False, attn_drop: float = 0.0, max_num_nodes=None):
    """DOSTransformer.forward -- embedder_eDOS/DOSTransformer.py:45-93.
    Returns (dos_global [B,T], x [N,H], dos_system [B,T])."""
    L = len({k.split(".")[1] for k in p if k.startswith("stacked_processor.")})
    t = len({k.split(".")[2] for k in p if k.startswith("transformer.layers.")})
    x = _mlp_prelu(p, "GN_encoder.node_encoder", g.x)
    e = _mlp_prelu(p, "GN_encoder.edge_encoder", g.edge_attr)
    u = _mlp_prelu(p, "GN_encoder.global_encoder", g.glob.reshape(-1, 2))
    row, col = g.edge_index[0], g.edge_index[1]
    x, e = _message_passing(p, x, e, row, col, L, mean=False)
    B = u.shape[0]
    pooled = segment_sum(x, g.batch, B)
    graph = _lin(p, "GN_decoder.mlp.0", torch.cat([u, pooled], dim=1))       # Decoder :156-161
    dg, ds = _dos_heads(p, x, g.batch, g.system, p["embeddings.weight"], graph, "promt_token", t,
                        attn_drop, training, max_num_nodes)
    return dg, x, ds


def phonon_forward(p: Params, g, *, training: bool = False, attn_drop: float = 0.0, max_num_nodes=None):
    """DOSTransformer_phonon.forward -- embedder_phDOS/DOSTransformer_phonon.py:66-119."""
    L = len({k.split(".")[1] for k in p if k.startswith("stacked_processor.")})
    t = len({k.split(".")[2] for k in p if k.startswith("transformer.layers.")})
    ea = phonon_edge_features(g["edge_vec"])
    x = _mlp_prelu(p, "GN_encoder.node_encoder", g.x)
    e = _mlp_prelu(p, "GN_encoder.edge_encoder", ea)
    row, col = g.edge_index[0], g.edge_index[1]
    x, e = _message_passing(p, x, e, row, col, L, mean=True)
    B = int(g.batch.max()) + 1
    graph = _lin(p, "GN_decoder.mlp.0", segment_sum(x, g.batch, B))          # Decoder :178-183
    dg, ds = _dos_heads(p, x, g.batch, g.system, p["embeddings.weight"], graph, "prompt_token", t,
                        attn_drop, training, max_num_nodes)
    return dg, x, ds


# ------------------------------------------------------------------ losses / metrics
def edos_loss(dos_global, dos_system, y_ft, beta: float = 1.0):
    """main_eDOS.py:111-123: clamp targets at 0, per-crystal RMSE, mean over crystals, + beta * system."""
    B = dos_global.shape[0]
    y = torch.clamp(y_ft, min=0).reshape(B, -1)
    rg = torch.sqrt(((y - dos_global) ** 2).mean(dim=1)).mean()
    rs = torch.sqrt(((y - dos_system) ** 2).mean(dim=1)).mean()
    return rg + beta * rs


def phonon_loss(dos_global, dos_system, phdos, beta: float = 1.0):
    """main_phDOS.py:109-114: sqrt of the batch-wide MSE, + beta * system."""
    y = phdos.reshape(dos_global.shape[0], -1)
    return torch.sqrt(F.mse_loss(dos_global, y)) + beta * torch.sqrt(F.mse_loss(dos_system, y))


def eval_metrics(dos_system, target, clamp_pred: bool) -> Dict[str, torch.Tensor]:
    """utils.py:74-88 (eDOS, clamp_pred=True; targets clamped too) / :127-138 (phonon):
    per-batch mean MSE, RMSE, MAE and variance-weighted R^2."""
    B = dos_system.shape[0]
    y = target.reshape(B, -1)
    if clamp_pred:
        y = torch.clamp(y, min=0)
        dos_system = torch.clamp(dos_system, min=0)
    mse = ((y - dos_system) ** 2).mean(dim=1)
    yf, pf = y.flatten(), dos_system.flatten()
    r2 = 1.0 - ((yf - pf) ** 2).sum() / ((yf - yf.mean()) ** 2).sum()
    return {"mse": mse.mean(), "rmse": torch.sqrt(mse).mean(), "mae": (y - dos_system).abs().mean(), "r2": r2}


# ------------------------------------------------------------------ integer structures (bit-exact spec)
def csr_by_key(key: torch.Tensor, size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stable sort of edge ids by ``key`` -> (rowptr [size+1], perm [E]).  This is
    the order torch_scatter's CPU loop / index_add_ visits the edges of one
    destination (ascending edge id), i.e. the summation order of the reference."""
    perm = torch.sort(key, stable=True).indices
    rowptr = torch.cat([key.new_zeros(1), torch.bincount(key, minlength=size).cumsum(0)])
    return rowptr, perm


def crystal_ptr(batch: torch.Tensor, B: Optional[int] = None) -> Tuple[torch.Tensor, int]:
    B = int(batch.max()) + 1 if B is None else B
    n = torch.bincount(batch, minlength=B)
    return torch.cat([n.new_zeros(1), n.cumsum(0)]), int(n.max())


def collate(graphs) -> Dict[str, object]:
    """Restatement of the PyG collate the launchers run on the CPU (torch_geometric.loader.DataLoader ->
    Batch.from_data_list; call sites main_eDOS.py:54-56, main_phDOS.py:52-54; PyG is absent and unpinned, so this follows
    its published semantics): tensors are concatenated along dim 0 in list order (0-d tensors are stacked),
    ``edge_index`` along dim 1 after adding the running node offset, ``batch`` = repeat_interleave(arange(B), n),
    ``ptr`` = node offsets, non-tensor fields become lists.  ``graphs``: mappings with crystal-local ``edge_index``."""
    keys = list(graphs[0].keys())
    out: Dict[str, object] = {}
    n = torch.tensor([int(g["x"].shape[0]) for g in graphs], dtype=torch.int64)
    ptr = torch.cat([n.new_zeros(1), n.cumsum(0)])
    for k in keys:
        vals = [g[k] for g in graphs]
        if not torch.is_tensor(vals[0]):
            out[k] = list(vals)
        elif k == "edge_index":
            out[k] = torch.cat([v + ptr[i] for i, v in enumerate(vals)], dim=1)
        elif vals[0].dim() == 0:
            out[k] = torch.stack(vals)
        else:
            out[k] = torch.cat(vals, dim=0)
    out["batch"] = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.int64), n)
    out["ptr"] = ptr
    return out


def state_dict_of(module: torch.nn.Module, dtype=None) -> Dict[str, torch.Tensor]:
    sd = {k: v.detach().clone() for k, v in module.state_dict().items()}
    if dtype is not None:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    return sd


def run_train_step(forward, loss_fn, params: Dict[str, torch.Tensor], g, target, beta=1.0):
    """fwd + bwd through the oracle; returns (outputs, loss, grads by name)."""
    leaf = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in params.items()}
    dg, x, ds = forward(leaf, g, training=True)
    loss = loss_fn(dg, ds, target, beta)
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    return (dg.detach(), x.detach(), ds.detach()), loss.detach(), grads

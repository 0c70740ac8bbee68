"""Stand-ins for the un-vendored third-party primitives the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference model files import (embedder_eDOS/DOSTransformer.py:5-7,
embedder_phDOS/DOSTransformer_phonon.py:5-10):

  torch_geometric.utils.to_dense_batch      torch_scatter.scatter_sum / scatter_mean
  e3nn.o3.spherical_harmonics / o3.Irreps   e3nn.nn.models.gate_points_2101.smooth_cutoff
  torch_cluster.radius_graph  (import only; the branch that calls it is dead)

None of them is installed in this image and no version is pinned by the
reference (no requirements file).  The functions below restate the published
semantics of those libraries (torch_scatter 2.x, PyG 2.x, e3nn 0.5.x); parity
for them is "unpinned" by the reference, so these definitions are the spec.
``install()`` registers them in ``sys.modules`` so the reference's model code
imports and runs unchanged.
"""
from __future__ import annotations

import math
import sys
import types

import torch


# --------------------------------------------------------------------------- torch_scatter
def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_sum along dim 0 (the only form the reference uses:
    DOSTransformer.py:158,187; DOSTransformer_phonon.py:180; utils.py:91)."""
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return res.index_add_(0, index, src)


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_mean: sum / clamp(count, 1) (DOSTransformer_phonon.py:209)."""
    total = scatter_sum(src, index, dim, out, dim_size)
    cnt = scatter_sum(torch.ones(index.shape[0], dtype=src.dtype, device=src.device), index, 0, None,
                      total.shape[0])
    cnt = cnt.clamp(min=1)
    return total / cnt.view((-1,) + (1,) * (total.dim() - 1))


# --------------------------------------------------------------------------- torch_geometric
def to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    """torch_geometric.utils.to_dense_batch for a sorted ``batch`` vector
    (DOSTransformer.py:61, DOSTransformer_phonon.py:86)."""
    if batch is None:
        mask = torch.ones(1, x.size(0), dtype=torch.bool, device=x.device)
        return x.unsqueeze(0), mask
    if batch_size is None:
        batch_size = int(batch.max()) + 1
    num_nodes = torch.bincount(batch, minlength=batch_size)
    cum = torch.cat([num_nodes.new_zeros(1), num_nodes.cumsum(0)])
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max())
    pos = torch.arange(batch.size(0), device=x.device) - cum[batch] + batch * max_num_nodes
    out = x.new_full((batch_size * max_num_nodes,) + tuple(x.shape[1:]), fill_value)
    out[pos] = x
    mask = torch.zeros(batch_size * max_num_nodes, dtype=torch.bool, device=x.device)
    mask[pos] = True
    return out.view((batch_size, max_num_nodes) + tuple(x.shape[1:])), mask.view(batch_size, max_num_nodes)


# --------------------------------------------------------------------------- e3nn
def smooth_cutoff(x):
    """e3nn.nn.models.gate_points_2101.smooth_cutoff."""
    u = 2 * (x - 1)
    y = (math.pi * u).cos().neg().add(1).div(2)
    y[u > 0] = 0
    y[u < -1] = 1
    return y


class _Irreps:
    """Only ``Irreps.spherical_harmonics(lmax)`` is used, as an opaque token."""

    def __init__(self, lmax):
        self.lmax = lmax

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return _Irreps(lmax)


def spherical_harmonics(l, x, normalize, normalization="integral"):
    """e3nn.o3.spherical_harmonics restricted to what the reference calls:
    irreps '0e + 1o', normalize=True, normalization='component'
    (DOSTransformer_phonon.py:75).  Y_0 = 1, Y_1 = sqrt(3) * x/|x| in e3nn's
    (x, y, z) ordering for l=1 under 'component' normalisation; a zero vector
    normalises to zero (torch.nn.functional.normalize semantics)."""
    lmax = l.lmax if isinstance(l, _Irreps) else int(l)
    assert lmax == 1 and normalization == "component"
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)
    one = torch.ones_like(x[..., :1])
    return torch.cat([one, math.sqrt(3.0) * x], dim=-1)


def radius_graph(*args, **kwargs):  # pragma: no cover - dead branch in the reference
    raise NotImplementedError("torch_cluster.radius_graph is only reachable through a dead branch "
                              "(DOSTransformer_phonon.py:59 reads a non-existent self.max_radius)")


def install():
    """Register the stand-ins in sys.modules (idempotent)."""
    if "torch_scatter" in sys.modules and getattr(sys.modules["torch_scatter"], "_dost_shim", False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m._dost_shim = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("torch_scatter", scatter_sum=scatter_sum, scatter_mean=scatter_mean)
    tg_utils = mod("torch_geometric.utils", to_dense_batch=to_dense_batch)
    mod("torch_geometric", utils=tg_utils)
    mod("torch_cluster", radius_graph=radius_graph)
    o3 = mod("e3nn.o3", spherical_harmonics=spherical_harmonics, Irreps=_Irreps)
    gp = mod("e3nn.nn.models.gate_points_2101", smooth_cutoff=smooth_cutoff)
    models = mod("e3nn.nn.models", gate_points_2101=gp)
    enn = mod("e3nn.nn", models=models)
    mod("e3nn", o3=o3, nn=enn)

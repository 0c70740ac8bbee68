"""On-device batch assembly from a packed crystal store.

The reference launchers collate on the CPU: ``torch_geometric.loader.DataLoader(dataset, batch_size, shuffle)``
(main_eDOS.py:54-56, main_phDOS.py:52-54) calls ``Batch.from_data_list`` for every step and the batch is then copied
to the device field by field (main_eDOS.py:106).  Here the whole dataset is packed ONCE into flat tables resident in
HBM (`PackedCrystals`), and a training / inference step only sends the crystal ids of its batch (8 bytes per
crystal); the kernels of csrc/collate.cu assemble the batch there:

* node / edge tables (``x``, ``edge_attr``, ``edge_vec`` ...): segmented row copies in batch order;
* ``edge_index``: crystal-local ids + the crystal's node offset in the batch;
* ``batch`` = repeat_interleave(arange(B), n_b), ``ptr`` = node offsets;
* per-crystal fields (``glob`` [2] -> [2B], ``system`` () -> [B], ``y_ft`` [T] -> [B*T], ``phdos`` [1,T] -> [B,T]):
  row gathers, concatenated along dim 0 exactly as PyG does.

The sizes of the assembled tensors (N, E) and the ``to_dense_batch`` padding length come from a host copy of the
per-crystal counts, so nothing is read back from the device.  Bit-exact against oracle.dost_oracle.collate (tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .synthetic import CrystalBatch

_INDEX_FIELDS = ("edge_index",)
_SKIP_FIELDS = ("batch", "ptr", "max_num_nodes")


def _rows(t: torch.Tensor) -> int:
    return 1 if t.dim() == 0 else int(t.shape[0])


class PackedCrystals:
    """Flat, device-resident tables of a dataset of crystal graphs + the host metadata to size a batch."""

    def __init__(self, tables: Dict[str, torch.Tensor], kinds: Dict[str, str], inner: Dict[str, tuple],
                 scalar: Dict[str, bool], node_ptr: torch.Tensor, edge_ptr: torch.Tensor,
                 edge_index: Optional[torch.Tensor], host_lists: Dict[str, list], device):
        self.device = torch.device(device)
        self.tables = {k: v.to(self.device).contiguous() for k, v in tables.items()}
        self.kinds, self.inner, self.scalar = kinds, inner, scalar
        self.node_count = np.diff(node_ptr.numpy()).astype(np.int64)       # host copies: size a batch without a sync
        self.edge_count = np.diff(edge_ptr.numpy()).astype(np.int64)
        self.node_ptr = node_ptr.to(self.device)
        self.edge_ptr = edge_ptr.to(self.device)
        self.edge_index = None if edge_index is None else edge_index.to(self.device).contiguous()
        self.host_lists = host_lists
        self.num_crystals = int(self.node_count.shape[0])
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.device)

    def __len__(self) -> int:
        return self.num_crystals

    def nbytes(self) -> int:
        n = sum(t.numel() * t.element_size() for t in self.tables.values())
        if self.edge_index is not None:
            n += self.edge_index.numel() * 8
        return n + (self.node_ptr.numel() + self.edge_ptr.numel()) * 8

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_graphs(cls, graphs: Sequence, device="cuda") -> "PackedCrystals":
        """``graphs``: per-crystal objects with PyG ``Data``-style fields (attribute or item access through ``keys()``):
        ``x`` [n,F], ``edge_index`` [2,e] with crystal-local ids, edge tables [e,...], per-crystal tensors of a fixed
        shape, and non-tensor fields (``mp_id``) that are carried as host lists."""
        assert len(graphs) > 0, "empty dataset"
        keys = [k for k in _keys(graphs[0]) if k not in _SKIP_FIELDS]
        nn = np.array([int(_get(g, "x").shape[0]) for g in graphs], dtype=np.int64)
        has_edges = "edge_index" in keys
        ne = np.array([int(_get(g, "edge_index").shape[1]) for g in graphs], dtype=np.int64) if has_edges \
            else np.zeros(len(graphs), dtype=np.int64)
        tables, kinds, inner, scalar, host_lists = {}, {}, {}, {}, {}
        edge_index = None
        for k in keys:
            vals = [_get(g, k) for g in graphs]
            if not torch.is_tensor(vals[0]):
                host_lists[k] = list(vals)
                continue
            if k in _INDEX_FIELDS:
                edge_index = torch.cat([v.to(torch.int64) for v in vals], dim=1)
                continue
            rows = np.array([_rows(v) for v in vals], dtype=np.int64)
            scalar[k] = vals[0].dim() == 0
            inner[k] = tuple(vals[0].shape[1:]) if vals[0].dim() > 0 else ()
            const = bool(np.all(rows == rows[0]))
            if k == "x" or (not const and np.array_equal(rows, nn)):
                kinds[k] = "node"
            elif has_edges and not const and np.array_equal(rows, ne):
                kinds[k] = "edge"
            elif const:
                kinds[k] = "crystal"          # same row count for every crystal: one fixed-size record each
            else:
                raise ValueError(f"field {k!r}: per-crystal row counts follow neither the atoms nor the edges")
            flat = torch.cat([v.reshape(1) if v.dim() == 0 else v for v in vals], dim=0)
            if kinds[k] == "crystal":
                flat = flat.reshape(len(graphs), -1)
                inner[k] = (int(rows[0]),) + inner[k]      # record shape; rows of one crystal stay together
            if flat.element_size() * max(1, int(np.prod(flat.shape[1:]))) % 4 != 0:
                raise ValueError(f"field {k!r}: rows must be a multiple of 4 bytes")
            tables[k] = flat
        node_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(nn)]).astype(np.int64))
        edge_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(ne)]).astype(np.int64))
        return cls(tables, kinds, inner, scalar, node_ptr, edge_ptr, edge_index, host_lists, device)

    @classmethod
    def from_batch(cls, batch, device="cuda") -> "PackedCrystals":
        """Pack an already collated (host) batch: the inverse of collate, used to turn the synthetic generators'
        output into a dataset."""
        return cls.from_graphs(split_batch(batch), device)

    def shard_ids(self, ids, world: int, T: int, hidden: int = 256):
        """Data parallel: split the global batch ``ids`` into ``world`` length-balanced parts with equal crystal counts
        (+-1) (dp.shard_crystals on the host copy of the counts).  Returns (parts, nmax, weights): each rank collates
        ``parts[rank]``, sets ``model.max_num_nodes = nmax`` (the global to_dense_batch padding length) and scales its
        loss by ``weights[rank]`` so that the SUM-reduced gradients equal the single-process ones."""
        from . import dp
        ids_t = torch.as_tensor(ids, dtype=torch.int64).cpu()
        ids_np = ids_t.numpy()
        cost = dp.crystal_cost(torch.from_numpy(self.node_count[ids_np]), torch.from_numpy(self.edge_count[ids_np]), T, hidden)
        bins = dp.shard_crystals(cost.tolist(), world)
        parts = [ids_t[torch.as_tensor(b, dtype=torch.int64)] for b in bins]
        nmax = int(self.node_count[ids_np].max()) if ids_np.size else 0
        weights = [dp.loss_weight(len(b), len(ids_np)) for b in bins]
        return parts, nmax, weights

    # ------------------------------------------------------------------ the hot call
    def collate(self, ids) -> CrystalBatch:
        """Assemble the batch of crystals ``ids`` (sequence / CPU int64 tensor, any order, repeats allowed) on the device."""
        lib = L.lib()
        ids_host = torch.as_tensor(ids, dtype=torch.int64)
        if ids_host.is_cuda:
            ids_dev, ids_host = ids_host, ids_host.cpu()          # device ids cost one read-back; prefer host ids
        else:
            ids_dev = ids_host.to(self.device, non_blocking=True)
        ids_np = ids_host.numpy()
        B = int(ids_np.shape[0])
        if B and (ids_np.min() < 0 or ids_np.max() >= self.num_crystals):
            raise IndexError(f"crystal id outside [0, {self.num_crystals})")
        N = int(self.node_count[ids_np].sum())
        E = int(self.edge_count[ids_np].sum())
        st = L.stream()
        dev = self.device
        node_ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        edge_ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        if B:
            L.check(lib.dost_collate_ptr(L.p(ids_dev), B, L.p(self.node_ptr), L.p(self.edge_ptr), self.num_crystals,
                                         L.p(node_ptr), L.p(edge_ptr), None, L.p(self._bad), st), "collate_ptr")
        out = {}
        for k, tab in self.tables.items():
            kind = self.kinds[k]
            row_bytes = tab.element_size() * int(np.prod(tab.shape[1:])) if tab.dim() > 1 else tab.element_size()
            if kind == "crystal":
                dst = torch.empty((B,) + tuple(tab.shape[1:]), dtype=tab.dtype, device=dev)
                L.check(lib.dost_collate_rows(L.p(tab), None, L.p(ids_dev), None, B, B, row_bytes, L.p(dst), st),
                        "collate_rows")
                rec = self.inner[k]
                out[k] = dst.reshape(B) if self.scalar[k] else dst.reshape((B * rec[0],) + tuple(rec[1:]))
            else:
                rows, sp, op = (N, self.node_ptr, node_ptr) if kind == "node" else (E, self.edge_ptr, edge_ptr)
                dst = torch.empty((rows,) + tuple(tab.shape[1:]), dtype=tab.dtype, device=dev)
                L.check(lib.dost_collate_rows(L.p(tab), L.p(sp), L.p(ids_dev), L.p(op), B, rows, row_bytes, L.p(dst), st),
                        "collate_rows")
                out[k] = dst
        bvec = torch.empty(N, dtype=torch.int64, device=dev)
        ei = None
        if self.edge_index is not None:
            ei = torch.empty(2, E, dtype=torch.int64, device=dev)
        L.check(lib.dost_collate_index(L.p(self.edge_index), 0 if self.edge_index is None else self.edge_index.shape[1],
                                       L.p(self.node_ptr), L.p(self.edge_ptr), L.p(ids_dev), L.p(node_ptr),
                                       L.p(edge_ptr), B, N, E, L.p(ei), L.p(bvec), st), "collate_index")
        if ei is not None:
            out["edge_index"] = ei
        out["batch"] = bvec
        out["ptr"] = node_ptr
        for k, vals in self.host_lists.items():
            out[k] = [vals[i] for i in ids_np]
        out["max_num_nodes"] = int(self.node_count[ids_np].max()) if B else 0
        return CrystalBatch(**out)


# ---------------------------------------------------------------------- helpers
def _keys(g) -> List[str]:
    k = g.keys() if callable(getattr(g, "keys", None)) else g.keys
    return list(k)


def _get(g, k):
    return g[k] if isinstance(g, dict) else getattr(g, k)


def split_batch(batch) -> List[CrystalBatch]:
    """Host-side inverse of the collate: per-crystal graphs (crystal-local ``edge_index``) of a collated batch whose
    layout is the one SURVEY.md section 8b lists (``glob`` [2B], ``y_ft`` [B*T], ``phdos`` [B,T], ``system`` [B])."""
    keys = _keys(batch)
    bvec = _get(batch, "batch").cpu()
    B = int(_get(batch, "system").shape[0])
    n = torch.bincount(bvec, minlength=B)
    noff = torch.cat([n.new_zeros(1), n.cumsum(0)])
    N = int(noff[-1])
    ei = _get(batch, "edge_index").cpu() if "edge_index" in keys else None
    if ei is not None:
        eb = bvec[ei[0]]
        assert bool((eb[1:] >= eb[:-1]).all()), "edges must be grouped by crystal"
        e = torch.bincount(eb, minlength=B)
        eoff = torch.cat([e.new_zeros(1), e.cumsum(0)])
        E = int(eoff[-1])
    graphs = []
    for b in range(B):
        g = {}
        for k in keys:
            if k in _SKIP_FIELDS:
                continue
            v = _get(batch, k)
            if not torch.is_tensor(v):
                g[k] = v[b] if isinstance(v, (list, tuple)) and len(v) == B else v
                continue
            v = v.cpu()
            if k == "edge_index":
                g[k] = v[:, eoff[b]:eoff[b + 1]] - noff[b]
            elif k == "x" or (v.shape[0] == N and N != B):
                g[k] = v[noff[b]:noff[b + 1]]
            elif ei is not None and v.shape[0] == E and E != B:
                g[k] = v[eoff[b]:eoff[b + 1]]
            elif v.shape[0] == B and v.dim() == 1:
                g[k] = v[b]                                   # 0-d per crystal (``system``)
            elif v.shape[0] == B:
                g[k] = v[b:b + 1]                             # [1, T] per crystal (``phdos``)
            else:
                assert v.shape[0] % B == 0, f"cannot split field {k!r}"
                r = v.shape[0] // B
                g[k] = v[b * r:(b + 1) * r]                   # ``glob`` [2], ``y_ft`` [T]
        graphs.append(CrystalBatch(**g))
    return graphs

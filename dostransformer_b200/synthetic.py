"""Synthetic crystal-graph batches of the reference's eDOS / phonon shapes.

The reference trains on PyG ``Batch`` objects built offline by
``data/mat2graph.py`` (eDOS) and ``utils.build_data`` (phonon).  Neither PyG
nor the datasets are available here, so the tests and ``bench.py`` use
seeded synthetic crystals with the same field names, dtypes and layout
(SURVEY.md section 8b / 8d):

eDOS   (mat2graph.py:146-159,212-243): ``x`` [N,200] float (last node of every
       crystal is the all-zero "prompt" node with no edges), ``edge_index``
       [2,E] int64 with row = centre atom / col = neighbour, K=12 out-edges per
       real atom, ``edge_attr`` [E,41] Gaussian distance expansion,
       ``glob`` [2B], ``batch`` [N] sorted, ``system`` [B] in 0..6,
       ``y_ft`` [B*201], ``mp_id`` list[str].
phonon (utils.py:249-303): ``x`` [N,118] mass-weighted one-hot, ``edge_index``
       with self-interaction edges, ``edge_vec`` [E,3], ``batch``, ``system``,
       ``phdos`` [B,51].
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

EDOS_T = 201
PHONON_T = 51


class CrystalBatch:
    """Duck-typed stand-in for a PyG ``Batch`` (attribute access, ``in``, ``[]``, ``.to``)."""

    def __init__(self, **fields):
        self._keys: List[str] = []
        for k, v in fields.items():
            setattr(self, k, v)
            self._keys.append(k)

    def keys(self):
        return list(self._keys)

    def __contains__(self, key):
        return key in self._keys

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        if key not in self._keys:
            self._keys.append(key)
        setattr(self, key, value)

    def _map(self, fn):
        for k in self._keys:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, fn(v))
        return self

    def to(self, device, non_blocking: bool = False):
        """In place, like PyG's ``Batch.to`` as the launchers use it (main_eDOS.py:106)."""
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def clone(self):
        def cp(v):
            if torch.is_tensor(v):
                return v.clone()
            return list(v) if isinstance(v, (list, tuple)) else v
        return CrystalBatch(**{k: cp(getattr(self, k)) for k in self._keys})

    @property
    def num_graphs(self) -> int:
        return int(self.system.shape[0])

    def nbytes(self) -> int:
        return sum(getattr(self, k).numel() * getattr(self, k).element_size()
                   for k in self._keys if torch.is_tensor(getattr(self, k)))


def _lognormal_sizes(B, gen, mean_atoms, sigma, lo, hi):
    z = torch.randn(B, generator=gen, dtype=torch.float64)
    n = torch.round(torch.exp(math.log(mean_atoms) + sigma * z)).clamp_(lo, hi).to(torch.int64)
    return n


def make_edos_batch(B: int, seed: int = 2000, *, mean_atoms: float = 20.0, sigma: float = 0.6,
                    min_atoms: int = 1, max_atoms: int = 200, K: int = 12, n_atom_feats: int = 200,
                    T: int = EDOS_T, sizes: Optional[torch.Tensor] = None,
                    dtype: torch.dtype = torch.float32) -> CrystalBatch:
    """eDOS random-split shape."""
    gen = torch.Generator().manual_seed(int(seed))
    n_real = sizes.to(torch.int64) if sizes is not None else _lognormal_sizes(B, gen, mean_atoms, sigma,
                                                                             min_atoms, max_atoms)
    n_nodes = n_real + 1                                     # + the zero-feature node (mat2graph.py:155-158)
    node_off = torch.cat([torch.zeros(1, dtype=torch.int64), n_nodes.cumsum(0)])
    N = int(node_off[-1])
    batch = torch.repeat_interleave(torch.arange(B, dtype=torch.int64), n_nodes)
    is_real = torch.ones(N, dtype=torch.bool)
    is_real[node_off[1:] - 1] = False
    x = torch.randn(N, n_atom_feats, generator=gen, dtype=torch.float64)
    x[~is_real] = 0.0
    real_idx = torch.nonzero(is_real).squeeze(1)             # global ids of real atoms, ascending
    row = real_idx.repeat_interleave(K)                      # centre atom (mat2graph.py:239)
    E = row.numel()
    crystal_of_edge = batch[row]
    u = torch.rand(E, generator=gen, dtype=torch.float64)
    col_local = torch.floor(u * n_real[crystal_of_edge].to(torch.float64)).to(torch.int64)
    col_local = torch.minimum(col_local, n_real[crystal_of_edge] - 1)
    col = node_off[crystal_of_edge] + col_local
    d = 0.8 + 7.2 * torch.rand(real_idx.numel(), K, generator=gen, dtype=torch.float64)
    d = d.sort(dim=1).values.reshape(-1)                     # neighbours sorted by distance
    centres = torch.arange(0.0, 8.0 + 1e-9, 0.2, dtype=torch.float64)   # 41 centres (mat2graph.py:162-179)
    edge_attr = torch.exp(-((d[:, None] - centres[None, :]) ** 2) / (0.2 ** 2))
    glob = torch.randn(2 * B, generator=gen, dtype=torch.float64)
    system = torch.randint(0, 7, (B,), generator=gen, dtype=torch.int64)
    y = torch.randn(B, T + 4, generator=gen, dtype=torch.float64).abs()
    y = y.unfold(1, 5, 1).mean(-1)                           # light smoothing -> [B, T]
    y = y / y.amax(dim=1, keepdim=True)
    neg = torch.rand(B, T, generator=gen, dtype=torch.float64) < 0.05   # exercise the clamp (main_eDOS.py:112)
    y = torch.where(neg, -0.05 * y, y)
    return CrystalBatch(
        x=x.to(dtype), edge_index=torch.stack([row, col]), edge_attr=edge_attr.to(dtype),
        glob=glob.to(dtype), batch=batch, system=system, y_ft=y.reshape(-1).to(dtype),
        mp_id=[f"syn-{seed}-{i}" for i in range(B)],
        # host-side collate metadata (not a reference field): the to_dense_batch padding length, known for free while the
        # batch is assembled on the CPU; lets the model size its ragged attention tiles without a device->host read
        max_num_nodes=int(n_nodes.max()))


def make_large_cell_batch(B: int, seed: int = 4000, *, K: int = 24, lo: int = 200, hi: int = 400,
                          T: int = EDOS_T, dtype: torch.dtype = torch.float32) -> CrystalBatch:
    """200-400 atoms per crystal, 24 neighbours."""
    gen = torch.Generator().manual_seed(int(seed) + 17)
    sizes = torch.randint(lo, hi + 1, (B,), generator=gen, dtype=torch.int64)
    return make_edos_batch(B, seed, K=K, sizes=sizes, T=T, dtype=dtype)


def make_phonon_batch(B: int = 1, seed: int = 1000, *, K: int = 24, min_atoms: int = 2, max_atoms: int = 12,
                      n_atom_feats: int = 118, T: int = PHONON_T,
                      dtype: torch.dtype = torch.float64) -> CrystalBatch:
    """phonon-DOS shape (utils.py:249-303)."""
    gen = torch.Generator().manual_seed(int(seed))
    n = torch.randint(min_atoms, max_atoms + 1, (B,), generator=gen, dtype=torch.int64)
    off = torch.cat([torch.zeros(1, dtype=torch.int64), n.cumsum(0)])
    N = int(off[-1])
    batch = torch.repeat_interleave(torch.arange(B, dtype=torch.int64), n)
    species = torch.randint(0, n_atom_feats, (N,), generator=gen, dtype=torch.int64)
    mass = 1.0 + 239.0 * torch.rand(N, generator=gen, dtype=torch.float64)
    x = torch.zeros(N, n_atom_feats, dtype=torch.float64)
    x[torch.arange(N), species] = mass
    src = torch.arange(N, dtype=torch.int64).repeat_interleave(K)
    E = src.numel()
    cb = batch[src]
    u = torch.rand(E, generator=gen, dtype=torch.float64)
    dst_local = torch.minimum(torch.floor(u * n[cb].to(torch.float64)).to(torch.int64), n[cb] - 1)
    dst = off[cb] + dst_local
    r = 0.5 + 3.5 * torch.rand(E, generator=gen, dtype=torch.float64)
    v = torch.randn(E, 3, generator=gen, dtype=torch.float64)
    v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-9) * r[:, None]
    first = torch.arange(N, dtype=torch.int64) * K            # first edge of each atom = self-interaction
    dst[first] = torch.arange(N, dtype=torch.int64)
    v[first] = 0.0
    system = torch.randint(0, 7, (B,), generator=gen, dtype=torch.int64)
    ph = torch.rand(B, T, generator=gen, dtype=torch.float64)
    ph = ph / ph.amax(dim=1, keepdim=True)
    return CrystalBatch(x=x.to(dtype), edge_index=torch.stack([src, dst]), edge_vec=v.to(dtype), batch=batch,
                        system=system, phdos=ph.to(dtype))


def pad_edos_batch(g: CrystalBatch, *, node_bucket: int = 256, dummies: int = 8, nmax_bucket: int = 32,
                   T: int = EDOS_T) -> CrystalBatch:
    """Pads an eDOS batch to bucketed shapes so that CUDA-graph replay (graphed.GraphedStep) sees few distinct signatures.

    ``dummies`` zero-feature dummy crystals are appended (each = some atoms + the zero node of mat2graph.py:155-158, every
    atom with the data set's fixed out-degree pointing at itself, zero targets) so that the node count becomes a multiple
    of ``node_bucket``; with a fixed out-degree the edge count follows.  ``n_valid`` = the number of real crystals: the
    loss is taken over the first ``n_valid`` rows only, so the dummies receive exactly zero gradient and contribute
    exactly zero to every weight gradient.  No dummy is larger than the largest real crystal, hence the padding length
    Nmax (= every real crystal's phantom-key count) is unchanged; the host-side hint ``max_num_nodes`` is
    rounded up to ``nmax_bucket`` (it only sizes buffers; the phantom count comes from the device-side maximum).
    CPU tensors in, CPU tensors out (a collate-time operation)."""
    B = int(g.system.numel())
    N = int(g.batch.numel())
    n_nodes = torch.bincount(g.batch, minlength=B)
    nmax = int(n_nodes.max())
    E = int(g.edge_index.shape[1])
    n_atoms = N - B
    K = E // max(n_atoms, 1)
    if n_atoms <= 0 or K * n_atoms != E:
        raise ValueError("pad_edos_batch needs the eDOS layout: one zero node per crystal and a fixed out-degree per atom")
    D = int(dummies)
    while True:
        N_to = -(-(N + 2 * D) // node_bucket) * node_bucket
        P = N_to - N
        sizes = [P // D + (1 if j < P % D else 0) for j in range(D)]
        if max(sizes) <= nmax or nmax < 2:
            break
        D *= 2
    if nmax < 2:
        raise ValueError("pad_edos_batch: the largest real crystal must have at least one atom")
    xs, rows, batches = [g.x], [], [g.batch]
    off = N
    for j, s in enumerate(sizes):
        atoms = torch.arange(off, off + s - 1, dtype=torch.int64)
        rows.append(atoms.repeat_interleave(K))
        batches.append(torch.full((s,), B + j, dtype=torch.int64))
        off += s
    row_pad = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.int64)
    xs.append(torch.zeros(P, g.x.shape[1], dtype=g.x.dtype))
    hint = -(-(nmax + 1) // nmax_bucket) * nmax_bucket - 1
    out = CrystalBatch(
        x=torch.cat(xs), edge_index=torch.cat([g.edge_index, torch.stack([row_pad, row_pad])], dim=1),
        edge_attr=torch.cat([g.edge_attr, torch.zeros(row_pad.numel(), g.edge_attr.shape[1], dtype=g.edge_attr.dtype)]),
        glob=torch.cat([g.glob, torch.zeros(2 * D, dtype=g.glob.dtype)]), batch=torch.cat(batches),
        system=torch.cat([g.system, torch.zeros(D, dtype=g.system.dtype)]),
        y_ft=torch.cat([g.y_ft, torch.zeros(D * T, dtype=g.y_ft.dtype)]),
        mp_id=list(getattr(g, "mp_id", [f"c{i}" for i in range(B)])) + [f"pad-{j}" for j in range(D)],
        max_num_nodes=hint, n_valid=B)
    return out

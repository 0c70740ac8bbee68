"""Graph construction on the device: periodic neighbour lists, the phonon edge vectors
and the eDOS 12-nearest-neighbour bonds with their Gaussian distance features.

The reference builds its graphs offline on the CPU with third-party code: ``ase.neighbor_list("ijS", cutoff=r_max,
self_interaction=True)`` + ``edge_vec = pos[dst] - pos[src] + shift @ lattice`` for the phonon data (utils.py:267-273),
``Structure.get_all_neighbors(radius)`` -> 12 nearest -> ``GaussianDistance.expand`` for the eDOS data
(data/mat2graph.py:162-179,185,212-243).  These functions take a whole batch / dataset of crystals at once
(``lattice`` [C,3,3], Cartesian ``pos`` [N,3], ``node_ptr`` [C+1]) and run the kernels of csrc/neighbors.cu; the edge order
is the library's canonical one (centre atom, neighbour atom, shift), see that file.  fp64 like the reference's geometry.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib as L


def _prep(lattice: torch.Tensor, pos: torch.Tensor, node_ptr: torch.Tensor):
    assert lattice.dim() == 3 and lattice.shape[1:] == (3, 3) and pos.dim() == 2 and pos.shape[1] == 3
    lattice = lattice.to(torch.float64).contiguous()
    pos = pos.to(torch.float64).contiguous()
    node_ptr = node_ptr.to(device=pos.device, dtype=torch.int64).contiguous()
    C = lattice.shape[0]
    assert node_ptr.numel() == C + 1
    counts = node_ptr[1:] - node_ptr[:-1]
    crystal_of = torch.repeat_interleave(torch.arange(C, device=pos.device, dtype=torch.int64), counts,
                                         output_size=pos.shape[0])
    return lattice, pos, node_ptr, crystal_of


def neighbor_list(lattice: torch.Tensor, pos: torch.Tensor, node_ptr: torch.Tensor, cutoff: float,
                  self_interaction: bool = True, local_ids: bool = False) -> Dict[str, torch.Tensor]:
    """All periodic images within ``cutoff`` of every atom.  Returns ``edge_index`` [2,E] (row 0 = centre, row 1 =
    neighbour; ids local to the crystal if ``local_ids`` else rows of ``pos``), ``edge_shift`` [E,3] int64, ``edge_vec`` [E,3],
    ``edge_len`` [E] and ``edge_ptr`` [N+1] (edges of atom i are edge_ptr[i]..edge_ptr[i+1]).  One device->host read
    (the edge total) sizes the outputs; this is the offline stage of the reference."""
    lib = L.lib()
    lattice, pos, node_ptr, crystal_of = _prep(lattice, pos, node_ptr)
    N = pos.shape[0]
    dev = pos.device
    st = L.stream()
    count = torch.zeros(N, dtype=torch.int64, device=dev)
    L.check(lib.dost_neighbor_count(L.p(lattice), L.p(pos), L.p(node_ptr), L.p(crystal_of), N, float(cutoff),
                                    int(self_interaction), L.p(count), st), "neighbor_count")
    edge_ptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(count, 0, out=edge_ptr[1:])
    E = int(edge_ptr[-1].item()) if N else 0
    ei = torch.empty(2, E, dtype=torch.int64, device=dev)
    shift = torch.empty(E, 3, dtype=torch.int64, device=dev)
    vec = torch.empty(E, 3, dtype=torch.float64, device=dev)
    length = torch.empty(E, dtype=torch.float64, device=dev)
    if E:
        L.check(lib.dost_neighbor_fill(L.p(lattice), L.p(pos), L.p(node_ptr), L.p(crystal_of), N, float(cutoff),
                                       int(self_interaction), L.p(edge_ptr), int(local_ids), L.p(ei[0]), L.p(ei[1]),
                                       L.p(shift), L.p(vec), L.p(length), st), "neighbor_fill")
    return {"edge_index": ei, "edge_shift": shift, "edge_vec": vec, "edge_len": length, "edge_ptr": edge_ptr}


def phonon_edges(lattice, pos, node_ptr, r_max: float = 4.0) -> Dict[str, torch.Tensor]:
    """utils.build_data's edges (utils.py:267-273) for a batch of crystals: self-interaction images included."""
    return neighbor_list(lattice, pos, node_ptr, r_max, self_interaction=True, local_ids=False)


def knn_select(nl: Dict[str, torch.Tensor], k: int, pad_idx: int, pad_dist: float) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = L.lib()
    edge_ptr = nl["edge_ptr"]
    N = edge_ptr.numel() - 1
    dev = edge_ptr.device
    idx = torch.empty(N, k, dtype=torch.int64, device=dev)
    dist = torch.empty(N, k, dtype=torch.float64, device=dev)
    dst = nl["edge_index"][1].contiguous()
    L.check(lib.dost_knn_select(L.p(edge_ptr), L.p(dst), L.p(nl["edge_len"]), N, k, pad_idx, float(pad_dist), L.p(idx),
                                L.p(dist), None, L.stream()), "knn_select")
    return idx, dist


def gaussian_expand(dist: torch.Tensor, dmin: float = 0.0, dmax: float = 8.0, step: float = 0.2) -> torch.Tensor:
    """GaussianDistance(dmin, dmax, step).expand + fp32 cast (mat2graph.py:162-179,235): [..., nfilt]."""
    lib = L.lib()
    import numpy as np
    nfilt = int(np.arange(dmin, dmax + step, step).shape[0])      # the reference's own filter count (41 for 0..8 by 0.2)
    flat = dist.to(torch.float64).contiguous().reshape(-1)
    out = torch.empty(flat.numel(), nfilt, dtype=torch.float32, device=dist.device)
    L.check(lib.dost_gaussian_expand(L.p(flat), flat.numel(), float(dmin), float(step), nfilt, float(step), L.p(out),
                                     L.stream()), "gaussian_expand")
    return out.reshape(tuple(dist.shape) + (nfilt,))


def edos_edges(lattice, pos, node_ptr, radius: float = 8.0, k: int = 12) -> Dict[str, torch.Tensor]:
    """get_bond_info (mat2graph.py:212-243) for a batch of crystals: ``edge_index`` [2, N*k] (row 0 = centre atom, row 1 =
    neighbour; rows of ``pos``, padded slots point at the crystal's atom 0), ``edge_attr`` [N*k, 41] fp32,
    ``nbr_dist`` [N,k]."""
    nl = neighbor_list(lattice, pos, node_ptr, radius, self_interaction=False, local_ids=True)
    idx, dist = knn_select(nl, k, 0, radius + 1.0)
    N = pos.shape[0]
    node_ptr = node_ptr.to(device=pos.device, dtype=torch.int64)
    counts = node_ptr[1:] - node_ptr[:-1]
    first = torch.repeat_interleave(node_ptr[:-1], counts, output_size=N)           # first row of each atom's crystal
    centre = torch.arange(N, device=pos.device, dtype=torch.int64).repeat_interleave(k)
    nbr = (idx + first[:, None]).reshape(-1)
    feats = gaussian_expand(dist, 0.0, radius, 0.2)
    return {"edge_index": torch.stack([centre, nbr]), "edge_attr": feats.reshape(N * k, -1), "nbr_dist": dist,
            "nbr_idx_local": idx}


def edos_graph_batch(lattice, pos, node_ptr, atom_feats: torch.Tensor, radius: float = 8.0, k: int = 12, **per_crystal):
    """Structures -> the eDOS model's input batch, on the device: what ``get_crystal_graph`` (mat2graph.py:120-159) builds
    per crystal and PyG then collates.  ``atom_feats`` [N,F] are the (already scaled) element feature rows of the atoms; every
    crystal gets the all-zero prompt node appended after its atoms (mat2graph.py:155-158), which has no bonds; bonds are the
    k nearest images within ``radius`` with their Gaussian distance features.  ``per_crystal`` tensors (``glob`` [2C],
    ``system`` [C], ``y_ft`` [C*T] ...) are passed through."""
    from .synthetic import CrystalBatch
    e = edos_edges(lattice, pos, node_ptr, radius, k)
    dev = pos.device
    node_ptr = node_ptr.to(device=dev, dtype=torch.int64)
    C = node_ptr.numel() - 1
    N = pos.shape[0]
    counts = node_ptr[1:] - node_ptr[:-1]
    crystal_of = torch.repeat_interleave(torch.arange(C, device=dev, dtype=torch.int64), counts, output_size=N)
    x = torch.zeros(N + C, atom_feats.shape[1], dtype=atom_feats.dtype, device=dev)
    rows = torch.arange(N, device=dev, dtype=torch.int64) + crystal_of        # one prompt node per preceding crystal
    x[rows] = atom_feats.to(dev)
    ei = e["edge_index"]
    edge_index = torch.stack([ei[0] + crystal_of[ei[0]], ei[1] + crystal_of[ei[1]]])
    batch = torch.repeat_interleave(torch.arange(C, device=dev, dtype=torch.int64), counts + 1, output_size=N + C)
    # one host read for the padding length; this is the offline stage (the training loop gets it from the packed store)
    nmax = int(counts.max().item()) + 1 if C else 0
    return CrystalBatch(x=x, edge_index=edge_index, edge_attr=e["edge_attr"], batch=batch, max_num_nodes=nmax,
                        **per_crystal)


def phonon_graph_batch(lattice, pos, node_ptr, x: torch.Tensor, r_max: float = 4.0, **per_crystal):
    """Structures -> the phonon model's input batch on the device (utils.build_data, utils.py:249-303, then the PyG
    collate): ``x`` [N,118] are the mass-weighted one-hot rows of the atoms (the element table stays on the host),
    edges = every periodic image within ``r_max`` including the self-interaction images, ``edge_vec`` = pos[dst] - pos[src]
    + shift @ lattice.  ``per_crystal`` tensors (``system`` [C], ``phdos`` [C,51] ...) are passed through."""
    from .synthetic import CrystalBatch
    nl = phonon_edges(lattice, pos, node_ptr, r_max)
    dev = pos.device
    node_ptr = node_ptr.to(device=dev, dtype=torch.int64)
    C = node_ptr.numel() - 1
    counts = node_ptr[1:] - node_ptr[:-1]
    batch = torch.repeat_interleave(torch.arange(C, device=dev, dtype=torch.int64), counts, output_size=pos.shape[0])
    return CrystalBatch(x=x.to(dev), edge_index=nl["edge_index"], edge_vec=nl["edge_vec"], edge_shift=nl["edge_shift"],
                        batch=batch, **per_crystal)

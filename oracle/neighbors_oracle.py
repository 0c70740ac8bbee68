"""CPU restatement (numpy, brute force) of the reference's offline graph construction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference delegates this to third-party code
that is neither vendored nor version-pinned (``ase.neighborlist.neighbor_list`` in utils.py:267,
``pymatgen.core.Structure.get_all_neighbors`` in data/mat2graph.py:185), and the reference has no fixtures for it.  What is
restated here is their published semantics -- every periodic image within the cutoff, self-interaction images included for
the phonon graphs -- on the reference's own call sites:

* phonon (utils.py:267-273): ``neighbor_list("ijS", cutoff=r_max, self_interaction=True)`` and
  ``edge_vec = pos[dst] - pos[src] + shift @ lattice``;
* eDOS (mat2graph.py:212-243): neighbours sorted by distance, the first 12 kept, short lists padded with index 0 and
  distance ``radius + 1``; bond features ``GaussianDistance(0, radius, 0.2).expand`` (mat2graph.py:162-179) cast to fp32.

The edge ORDER those libraries emit is an implementation detail of theirs; this oracle and the CUDA kernels
(csrc/neighbors.cu) share the canonical order (centre i, neighbour j, shift Sx, Sy, Sz) and the arithmetic contract written
in that file, so the comparison is bit-exact on indices, shifts, vectors and distances.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def _image_range(lattice: np.ndarray, pos: np.ndarray, cutoff: float) -> np.ndarray:
    """A generous bound on the lattice shifts that can bring any pair within the cutoff."""
    a, b, c = lattice
    vol = abs(np.dot(a, np.cross(b, c)))
    h = vol / np.array([np.linalg.norm(np.cross(b, c)), np.linalg.norm(np.cross(c, a)), np.linalg.norm(np.cross(a, b))])
    frac = pos @ np.linalg.inv(lattice)
    spread = frac.max(axis=0) - frac.min(axis=0) if len(pos) else np.zeros(3)
    return (np.ceil(cutoff / h + spread) + 1).astype(np.int64)


def neighbor_list(lattice: np.ndarray, pos: np.ndarray, cutoff: float, self_interaction: bool = True) -> Dict[str, np.ndarray]:
    """One crystal.  lattice [3,3] (rows = lattice vectors), pos [n,3] Cartesian, fp64.  Returns src, dst (local ids),
    shift [E,3] int64, vec [E,3], dist [E] in canonical order."""
    lattice = np.asarray(lattice, dtype=np.float64)
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[0]
    R = _image_range(lattice, pos, cutoff)
    sx, sy, sz = np.meshgrid(np.arange(-R[0], R[0] + 1), np.arange(-R[1], R[1] + 1), np.arange(-R[2], R[2] + 1), indexing="ij")
    S = np.stack([sx.ravel(), sy.ravel(), sz.ravel()], axis=1)                      # canonical (Sx, Sy, Sz) order
    Sf = S.astype(np.float64)
    # s_c = (Sx*L[0][c] + Sy*L[1][c]) + Sz*L[2][c], each operation rounded on its own (numpy ufuncs do not fuse)
    shift_vec = (Sf[:, 0:1] * lattice[0][None, :] + Sf[:, 1:2] * lattice[1][None, :]) + Sf[:, 2:3] * lattice[2][None, :]
    src, dst, shf, vec, dist = [], [], [], [], []
    for i in range(n):
        dp = pos - pos[i][None, :]                                                  # [n,3]: pos[j] - pos[i]
        v = dp[:, None, :] + shift_vec[None, :, :]                                  # [n, nS, 3]
        d = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2])
        ok = d < cutoff
        if not self_interaction:
            zero = (S == 0).all(axis=1)
            ok[i, zero] = False
        jj, ss = np.nonzero(ok)                                                     # row-major: j ascending, then shift order
        src.append(np.full(jj.shape, i, dtype=np.int64))
        dst.append(jj.astype(np.int64))
        shf.append(S[ss])
        vec.append(v[jj, ss])
        dist.append(d[jj, ss])
    cat = lambda xs, shape: np.concatenate(xs, axis=0) if xs else np.zeros(shape)
    return {"src": cat(src, (0,)).astype(np.int64), "dst": cat(dst, (0,)).astype(np.int64),
            "shift": cat(shf, (0, 3)).astype(np.int64), "vec": cat(vec, (0, 3)), "dist": cat(dist, (0,))}


def knn_from_list(nl: Dict[str, np.ndarray], n: int, k: int, radius: float) -> Tuple[np.ndarray, np.ndarray]:
    """mat2graph.py:217-231: per atom ``sorted(nbrs, key=distance)`` (stable), first k; short lists padded with index 0
    and distance radius + 1."""
    idx = np.zeros((n, k), dtype=np.int64)
    dist = np.full((n, k), radius + 1.0, dtype=np.float64)
    for i in range(n):
        m = np.nonzero(nl["src"] == i)[0]
        order = np.argsort(nl["dist"][m], kind="stable")[:k]
        idx[i, :len(order)] = nl["dst"][m][order]
        dist[i, :len(order)] = nl["dist"][m][order]
    return idx, dist


def gaussian_expand(dist: np.ndarray, dmin: float = 0.0, dmax: float = 8.0, step: float = 0.2) -> np.ndarray:
    """GaussianDistance(dmin, dmax, step).expand (mat2graph.py:162-179), then the fp32 cast of ``torch.Tensor(nbr_fea)``
    (mat2graph.py:235)."""
    filt = np.arange(dmin, dmax + step, step)
    var = step
    return np.exp(-(dist[..., np.newaxis] - filt) ** 2 / var ** 2).astype(np.float32)


def edos_edges(lattice: np.ndarray, pos: np.ndarray, radius: float = 8.0, k: int = 12) -> Tuple[np.ndarray, np.ndarray]:
    """get_bond_info (mat2graph.py:212-243) for one crystal: bonds [n*k, 2] (centre, neighbour) and bond_feats [n*k, 41]."""
    n = pos.shape[0]
    nl = neighbor_list(lattice, pos, radius, self_interaction=False)
    idx, dist = knn_from_list(nl, n, k, radius)
    feats = gaussian_expand(dist).reshape(-1, 41)
    centre = np.repeat(np.arange(n, dtype=np.int64), k)
    return np.stack([centre, idx.reshape(-1)], axis=1), feats


def random_crystal(rng: np.random.Generator, n: int, a_lo: float = 3.0, a_hi: float = 7.0) -> Tuple[np.ndarray, np.ndarray]:
    """A triclinic test cell (edge lengths in [a_lo, a_hi], mild shear) with n atoms at random fractional positions."""
    L = np.diag(rng.uniform(a_lo, a_hi, 3)) + rng.uniform(-0.8, 0.8, (3, 3)) * (1 - np.eye(3))
    frac = rng.uniform(0.0, 1.0, (n, 3))
    return L, frac @ L

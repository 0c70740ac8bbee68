"""GPU parity of the device graph construction (csrc/neighbors.cu) against the brute-force numpy oracle
(oracle/neighbors_oracle.py): indices, shifts, edge vectors and distances bit for bit; the fp32 Gaussian bond features to
at most one fp32 ulp on isolated elements (libm's and CUDA's fp64 exp may differ in the last bit before the cast)."""
import numpy as np
import pytest
import torch

from dostransformer_b200 import neighbors as NB
from oracle import neighbors_oracle as NO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _batch(seed, sizes, a_lo=3.0, a_hi=7.0):
    rng = np.random.default_rng(seed)
    cells = [NO.random_crystal(rng, n, a_lo, a_hi) for n in sizes]
    lattice = torch.tensor(np.stack([c[0] for c in cells]), dtype=torch.float64, device=DEV)
    pos = torch.tensor(np.concatenate([c[1] for c in cells]), dtype=torch.float64, device=DEV)
    node_ptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=DEV)
    return cells, lattice, pos, node_ptr


@pytest.mark.parametrize("self_interaction", [True, False])
def test_neighbor_list_bit_exact(self_interaction):
    sizes = [1, 2, 5, 17, 33, 3]
    cells, lattice, pos, node_ptr = _batch(11, sizes)
    got = NB.neighbor_list(lattice, pos, node_ptr, 4.0, self_interaction=self_interaction, local_ids=True)
    off = 0
    e0 = 0
    for (L, p), n in zip(cells, sizes):
        want = NO.neighbor_list(L, p, 4.0, self_interaction=self_interaction)
        E = len(want["dist"])
        sl = slice(e0, e0 + E)
        assert int(got["edge_ptr"][off + n]) - int(got["edge_ptr"][off]) == E
        assert np.array_equal(got["edge_index"][0, sl].cpu().numpy(), want["src"])
        assert np.array_equal(got["edge_index"][1, sl].cpu().numpy(), want["dst"])
        assert np.array_equal(got["edge_shift"][sl].cpu().numpy(), want["shift"])
        assert np.array_equal(got["edge_vec"][sl].cpu().numpy(), want["vec"])          # bit-exact fp64
        assert np.array_equal(got["edge_len"][sl].cpu().numpy(), want["dist"])
        off += n
        e0 += E
    assert e0 == got["edge_index"].shape[1]
    # global ids = local ids + crystal offset; the reference's edge_vec formula holds on the result (utils.py:271-273)
    glob = NB.neighbor_list(lattice, pos, node_ptr, 4.0, self_interaction=self_interaction, local_ids=False)
    cof = torch.repeat_interleave(torch.arange(len(sizes), device=DEV), node_ptr[1:] - node_ptr[:-1])
    src, dst = glob["edge_index"]
    assert torch.equal(src - node_ptr[cof[src]], got["edge_index"][0]) and torch.equal(cof[src], cof[dst])
    ref_vec = pos[dst] - pos[src] + torch.einsum("ni,nij->nj", glob["edge_shift"].double(), lattice[cof[src]])
    assert torch.allclose(ref_vec, glob["edge_vec"], rtol=0, atol=1e-12)
    if self_interaction:
        selfs = (src == dst) & (glob["edge_shift"] == 0).all(dim=1)
        assert int(selfs.sum()) == sum(sizes) and bool((glob["edge_len"][selfs] == 0).all())


def test_edos_edges_match_oracle():
    sizes = [4, 1, 9, 21]
    cells, lattice, pos, node_ptr = _batch(12, sizes, 3.5, 6.5)
    got = NB.edos_edges(lattice, pos, node_ptr, radius=8.0, k=12)
    off = 0
    bad = 0
    for (L, p), n in zip(cells, sizes):
        bonds, feats = NO.edos_edges(L, p, radius=8.0, k=12)
        sl = slice(off * 12, (off + n) * 12)
        assert np.array_equal(got["edge_index"][0, sl].cpu().numpy(), bonds[:, 0] + off)
        assert np.array_equal(got["edge_index"][1, sl].cpu().numpy(), bonds[:, 1] + off)
        g = got["edge_attr"][sl].cpu().numpy()
        assert g.dtype == np.float32 and g.shape == feats.shape
        ulp = np.abs(g.view(np.int32).astype(np.int64) - feats.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, ulp.max()
        bad += int((ulp > 0).sum())
        off += n
    assert bad <= 2, bad                               # expected 0; a last-bit exp difference can flip an fp32 rounding
    d = got["nbr_dist"]
    assert bool((d[:, 1:] >= d[:, :-1]).all()) and bool((d > 0).all())


def test_edos_edges_padding_and_empty():
    # sparse cells: fewer than 12 images within 8 A -> index 0 of the crystal and distance 9 (mat2graph.py:223-229)
    lattice = torch.tensor(np.stack([np.eye(3) * 7.5, np.eye(3) * 30.0]), dtype=torch.float64, device=DEV)
    pos = torch.tensor([[0.0, 0.0, 0.0], [3.0, 0.0, 0.0], [1.0, 1.0, 1.0]], dtype=torch.float64, device=DEV)
    node_ptr = torch.tensor([0, 2, 3], dtype=torch.int64, device=DEV)
    got = NB.edos_edges(lattice, pos, node_ptr)
    bonds, feats = NO.edos_edges(np.eye(3) * 7.5, pos[:2].cpu().numpy())
    assert np.array_equal(got["edge_index"][:, :24].cpu().numpy().T, bonds)
    assert np.array_equal(got["edge_attr"][:24].cpu().numpy(), feats)
    # the lone atom of the second crystal has no neighbour at all: 12 padded bonds to itself (its crystal's atom 0)
    assert bool((got["edge_index"][1, 24:] == 2).all()) and bool((got["nbr_dist"][2] == 9.0).all())
    # an empty batch is a no-op
    e = NB.neighbor_list(lattice[:0], pos[:0], torch.zeros(1, dtype=torch.int64, device=DEV), 4.0)
    assert e["edge_index"].shape == (2, 0) and e["edge_ptr"].tolist() == [0]


def test_edos_graph_batch_feeds_the_model():
    """Structures -> graph -> model on the device, against the oracle's per-crystal construction + PyG-style collate."""
    from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
    from oracle import dost_oracle as O
    sizes = [3, 7, 1, 12]
    cells, lattice, pos, node_ptr = _batch(13, sizes, 3.5, 6.0)
    gen = torch.Generator().manual_seed(3)
    feats = torch.randn(sum(sizes), 200, generator=gen)
    C = len(sizes)
    glob = torch.randn(2 * C, generator=gen)
    system = torch.randint(0, 7, (C,), generator=gen)
    got = NB.edos_graph_batch(lattice, pos, node_ptr, feats.to(DEV), glob=glob.to(DEV), system=system.to(DEV))
    graphs, off = [], 0
    for b, ((L, p), n) in enumerate(zip(cells, sizes)):
        bonds, bf = NO.edos_edges(L, p)
        graphs.append({"x": torch.cat([feats[off:off + n], torch.zeros(1, 200)]), "edge_index": torch.tensor(bonds.T),
                       "edge_attr": torch.tensor(bf), "glob": glob[2 * b:2 * b + 2], "system": system[b]})
        off += n
    want = O.collate(graphs)
    for k in ("x", "edge_index", "batch", "glob", "system"):
        assert torch.equal(got[k].cpu(), want[k]), k
    assert (got["edge_attr"].cpu().view(torch.int32) - want["edge_attr"].view(torch.int32)).abs().max() <= 1
    assert got.max_num_nodes == max(sizes) + 1
    torch.manual_seed(0)
    model = DOSTransformer(2, 1, 200, 41, 2, 128, torch.device(DEV), 0.0).to(DEV).eval()
    with torch.no_grad():
        dg, x, ds = model(got)
    assert dg.shape == (C, 201) and ds.shape == (C, 201) and x.shape == (sum(sizes) + C, 128)
    assert bool(torch.isfinite(dg).all()) and bool(torch.isfinite(ds).all())
    # same numbers as the CPU oracle of the model on the oracle-built batch (fp32 tolerance of the model tests)
    sd = O.state_dict_of(model.cpu())
    from dostransformer_b200.synthetic import CrystalBatch
    ref = O.edos_forward(sd, CrystalBatch(**{k: v for k, v in want.items() if k != "ptr"}))
    assert (dg.cpu() - ref[0]).abs().max() <= 1e-4 * ref[0].abs().max()


def test_phonon_graph_batch_feeds_the_model():
    """Structures -> phonon graph -> model on the device, against the oracle-built graph through the CPU oracle model."""
    from dostransformer_b200.embedder_phDOS.DOSTransformer_phonon import DOSTransformer_phonon
    from dostransformer_b200.synthetic import CrystalBatch
    from oracle import dost_oracle as O
    torch.set_default_dtype(torch.float64)
    sizes = [2, 5, 3]
    cells, lattice, pos, node_ptr = _batch(14, sizes, 3.0, 5.5)
    gen = torch.Generator().manual_seed(4)
    N, C = sum(sizes), len(sizes)
    x = torch.zeros(N, 118)
    x[torch.arange(N), torch.randint(0, 118, (N,), generator=gen)] = 1.0 + 200.0 * torch.rand(N, generator=gen)
    system = torch.randint(0, 7, (C,), generator=gen)
    got = NB.phonon_graph_batch(lattice, pos, node_ptr, x.to(DEV), r_max=4.0, system=system.to(DEV))
    graphs, off = [], 0
    for b, ((L, p), n) in enumerate(zip(cells, sizes)):
        nl = NO.neighbor_list(L, p, 4.0, self_interaction=True)
        graphs.append({"x": x[off:off + n], "edge_index": torch.tensor(np.stack([nl["src"], nl["dst"]])),
                       "edge_vec": torch.tensor(nl["vec"]), "system": system[b]})
        off += n
    want = O.collate(graphs)
    for k in ("x", "edge_index", "edge_vec", "batch", "system"):
        assert torch.equal(got[k].cpu(), want[k]), k
    torch.manual_seed(0)
    model = DOSTransformer_phonon(2, 1, 118, 4, 64, torch.device(DEV), 0.0).to(DEV).eval()
    with torch.no_grad():
        dg, xx, ds = model(got)
    assert dg.shape == (C, 51) and bool(torch.isfinite(dg).all()) and bool(torch.isfinite(ds).all())
    sd = O.state_dict_of(model.cpu())
    ref = O.phonon_forward(sd, CrystalBatch(**{k: v for k, v in want.items() if k != "ptr"}))
    assert (dg.cpu() - ref[0]).abs().max() <= 1e-5 * ref[0].abs().max()
    assert (ds.cpu() - ref[2]).abs().max() <= 1e-5 * ref[2].abs().max()

"""CPU tests of the data-parallel layer: sharder, batch slicing, and a world_size-2 gloo run of the bucketed
gradient all-reduce whose result must equal the single-process gradient on the concatenated batch.
The compute on each rank is the CPU oracle (the CUDA path has no CPU fallback); the DP layer is model-agnostic."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import relerr
from dostransformer_b200 import dp
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch
from oracle import dost_oracle as O


def test_shard_crystals_balanced_and_deterministic():
    cost = [float(c) for c in torch.rand(37, generator=torch.Generator().manual_seed(0)) * 10 + 1]
    bins = dp.shard_crystals(cost, 4)
    assert sorted(i for b in bins for i in b) == list(range(37))
    assert sorted(len(b) for b in bins) == [9, 9, 9, 10]
    loads = [sum(cost[i] for i in b) for b in bins]
    assert max(loads) / min(loads) < 1.15
    assert bins == dp.shard_crystals(cost, 4)
    assert dp.shard_crystals([1.0], 2) == [[0], []]


def test_take_crystals_reindexes_like_a_fresh_collate():
    g = make_edos_batch(7, seed=3, mean_atoms=5.0)
    ids = [5, 1, 2]
    sub = dp.take_crystals(g, ids)
    n = torch.bincount(g.batch)
    assert torch.equal(torch.bincount(sub.batch), n[ids])
    off = torch.cat([n.new_zeros(1), n.cumsum(0)])
    for j, b in enumerate(ids):
        sel = sub.batch == j
        assert torch.equal(sub.x[sel], g.x[off[b]:off[b + 1]])
    assert torch.all(sub.batch[sub.edge_index[0]] == sub.batch[sub.edge_index[1]])
    assert sub.edge_index.max() < sub.batch.numel() and sub.edge_attr.shape[0] == sub.edge_index.shape[1]
    assert torch.equal(sub.system, g.system[ids]) and sub.mp_id == [g.mp_id[i] for i in ids]
    assert torch.equal(sub.y_ft.view(3, -1), g.y_ft.view(7, -1)[ids])
    # the sub-batch evaluated with the global padding length reproduces the rows of the full batch
    torch.manual_seed(0)
    sd = O.state_dict_of(DOSTransformer(2, 1, 200, 41, 2, 32, "cpu", 0.0))
    full = O.edos_forward(sd, g)[0]
    part = O.edos_forward(sd, sub, max_num_nodes=int(n.max()))[0]
    assert relerr(part, full[ids]) < 1e-5


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    torch.manual_seed(0)
    model = DOSTransformer(2, 1, 200, 41, 2, 32, "cpu", 0.0)
    g = make_edos_batch(6, seed=11, mean_atoms=5.0)
    parts, nmax, weights, bins = dp.shard_batch(g, world, T=201, hidden=32)
    reducer = dp.GradReducer(dp.live_named_parameters(model), bucket_bytes=64 << 10)
    params = dict(model.named_parameters())
    params.update(dict(model.named_buffers()))
    mine = parts[rank]
    for _ in range(2):                      # two steps: the reducer must be reusable
        model.zero_grad(set_to_none=True)
        dg, x, ds = O.edos_forward(params, mine, training=True, max_num_nodes=nmax)
        loss = O.edos_loss(dg, ds, mine.y_ft) * weights[rank]
        loss.backward()
        reducer.finish()
    if rank == 0:
        torch.save({"grads": {k: p.grad.clone() for k, p in dp.live_named_parameters(model)},
                    "nbuckets": len(reducer.buckets)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    out_path = str(tmp_path / "rank0.pt")
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = torch.load(out_path)
    grads, nbuckets = res["grads"], res["nbuckets"]
    assert nbuckets > 1
    torch.manual_seed(0)
    model = DOSTransformer(2, 1, 200, 41, 2, 32, "cpu", 0.0)
    g = make_edos_batch(6, seed=11, mean_atoms=5.0)
    _, _, ref = O.run_train_step(O.edos_forward, O.edos_loss, O.state_dict_of(model), g, g.y_ft)
    live = {k for k, _ in dp.live_named_parameters(model)}
    assert live == set(ref)                                   # the bucketed set is exactly the live set
    for k, r in ref.items():
        assert relerr(grads[k], r) < 2e-4, k


def test_packed_store_shards_ids_like_shard_batch():
    """PackedCrystals.shard_ids (host metadata only) gives the same partition, padding length and loss weights as
    dp.shard_batch on the collated batch of the same crystals."""
    from dostransformer_b200.collate import PackedCrystals, split_batch
    g = make_edos_batch(23, seed=77)
    store = PackedCrystals.from_graphs(split_batch(g), device="cpu")
    ids = torch.arange(23)
    for world in (2, 3, 8):
        parts, nmax, weights = store.shard_ids(ids, world, T=201)
        ref_parts, ref_nmax, ref_w, _ = dp.shard_batch(g, world, T=201)
        assert nmax == ref_nmax and weights == pytest.approx(ref_w)
        assert sorted(torch.cat(parts).tolist()) == ids.tolist()
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        for p, r in zip(parts, ref_parts):
            assert int(store.node_count[p.numpy()].sum()) == r.batch.numel()
            assert int(store.edge_count[p.numpy()].sum()) == r.edge_index.shape[1]
    # a shuffled subset: ids keep their identity through the partition
    sub = torch.tensor([5, 19, 2, 11, 7])
    parts, nmax, weights = store.shard_ids(sub, 2, T=201)
    assert sorted(torch.cat(parts).tolist()) == sorted(sub.tolist()) and abs(sum(weights) - 1.0) < 1e-12
    assert nmax == int(store.node_count[sub.numpy()].max())

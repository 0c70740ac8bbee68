"""Worker of tests/test_gpu_dp.py, launched with `python -m torch.distributed.run --nproc-per-node 2`: one rank per GPU,
NCCL.  Global batch -> dp.shard_batch (LPT sharder) -> CUDA model per rank -> loss * B_local / B_global -> backward with
the bucketed, overlapped GradReducer (or the graph-replay step with its flat all-reduce) -> reduced gradients to disk."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dostransformer_b200 import dp, ops  # noqa: E402
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer  # noqa: E402
from dostransformer_b200.graphed import GraphedStep  # noqa: E402
from dostransformer_b200.synthetic import make_edos_batch  # noqa: E402


def main():
    out_dir, precision, hidden, B = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = DOSTransformer(2, 2, 200, 41, 2, hidden, dev, 0.0, precision=precision).to(dev).train()
    g = make_edos_batch(B, seed=11, mean_atoms=9.0)
    parts, nmax, weights, bins = dp.shard_batch(g, world, T=201, hidden=hidden)
    model.max_num_nodes = nmax                      # global to_dense_batch padding length (phantom-key count)
    mine = parts[rank].to(dev)
    mine.max_num_nodes = int(torch.bincount(parts[rank].batch).max())
    reducer = dp.GradReducer(dp.live_named_parameters(model), bucket_bytes=256 << 10)
    runs = []
    for _ in range(2):                              # twice: the reduced gradients must be bit-identical across runs
        model.zero_grad(set_to_none=True)
        dg, x, ds = model(mine)
        loss_t = ops.dos_loss(dg, ds, mine.y_ft, mode="edos", beta=1.0) * weights[rank]
        loss_t.backward()
        loss = float(loss_t)
        reducer.finish()
        torch.cuda.synchronize()
        runs.append({k: p.grad.detach().cpu().clone() for k, p in dp.live_named_parameters(model)})
    reducer.remove()
    del dg, x, ds, loss_t
    # the graph-replay step (flat all-reduce after the replay) must give the same reduced gradients
    step = GraphedStep(model, "edos", loss_weight=weights[rank], world=world)
    step(mine)
    step(mine)
    torch.cuda.synchronize()
    graphed = {k: p.grad.detach().cpu().clone() for k, p in dp.live_named_parameters(model)}
    # AdamW fused into the all-reduce epilogue: every bucket is updated straight from its reduced flat buffer
    from dostransformer_b200.optim import AdamW
    del step
    opt = AdamW(model.parameters(), lr=1e-3, weight_decay=1e-2)
    reducer2 = dp.GradReducer(dp.live_named_parameters(model), bucket_bytes=256 << 10)
    model.zero_grad(set_to_none=True)
    dg, x, ds = model(mine)
    (ops.dos_loss(dg, ds, mine.y_ft, mode="edos", beta=1.0) * weights[rank]).backward()
    reducer2.finish(optimizer=opt)
    torch.cuda.synchronize()
    stepped = {k: p.detach().cpu().clone() for k, p in dp.live_named_parameters(model)}
    reducer2.remove()
    lsum = torch.tensor([float(loss)], device=dev, dtype=torch.float64)
    dist.all_reduce(lsum)
    if rank == 0:
        torch.save({"runs": runs, "graphed": graphed, "stepped": stepped, "nbuckets": len(reducer.buckets), "loss": float(lsum.item()),
                    "bins": bins, "nmax": nmax}, os.path.join(out_dir, "rank0.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

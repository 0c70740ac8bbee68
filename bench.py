#!/usr/bin/env python
"""bench.py -- train-step crystals/sec (fwd+bwd) of the eDOS DOSTransformer on synthetic crystal graphs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl product|reference] [--batch B_PER_GPU]

Workload (BASELINE.json configs[1]/[2]): DOSTransformer(3 GNN, 2 transformer layers, hidden 256), eDOS random-split
shape (~20 atoms/crystal log-normal, 12 neighbours, 200-d node features, 41-d edge features,
T = 201), B crystals per GPU (default 512), data-parallel over crystals with weak scaling (per-GPU batch fixed).
A step = graph prep + forward + loss + backward (+ gradient all-reduce when N > 1); no optimizer, no data loading,
as the metric is defined.

One JSON line is printed by rank 0.  `value` is measured with the batches resident in HBM; `e2e` repeats the same
steps from pinned host batches (H2D inside the timed region, loss read back every step).  `--impl reference` times
the CPU oracle port of the reference (the reference itself is pure Python and cannot travel to the GPU box;
oracle/dost_oracle.py restates it and is pinned against it by tests/golden) on the host cores.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

PREC_DESC = {"fp32": "fp32 FMA pipe", "bf16x3": "tcgen05 bf16x3 error-compensated operand planes (hi*hi + hi*lo + lo*hi), fp32 accumulate: outputs/loss <= 1e-4, gradients <= 2e-3 rel-L2 vs the fp64 reference",
             "bf16": "tcgen05 bf16 operands, fp32 accumulate (tolerance 5e-2 on gradients)"}
METRIC = "train-step crystals/sec (fwd+bwd)"
UNIT = "crystals/s"
HIDDEN, GNN_LAYERS, T_LAYERS, T = 256, 3, 2, 201


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d["bf16_tflops"]), tensor_sustained=float(
            d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  The process is started BEFORE the warm-up steps: its
    NVML attach takes 0.1-2 s (longer on an 8-GPU box) and stalls CUDA work on the box while it lasts, which must not land
    inside the timed region; only the samples stamped inside the region (`mark()` .. `stop()`) are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.lines, self.proc, self.t0 = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_ready(self, timeout: float = 8.0):
        """Blocks until the first sample has arrived (the NVML attach is over) or the timeout passes."""
        t_end = time.time() + timeout
        while self.proc is not None and not self.lines and time.time() < t_end and self.proc.poll() is None:
            time.sleep(0.02)

    def mark(self):
        """The timed region starts now."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [ln for ts, ln in self.lines if t0 <= ts <= t1 + 0.12]
        if not inside and self.lines:                    # region shorter than one sampling period: the closest sample
            inside = [min(self.lines, key=lambda x: abs(x[0] - t1))[1]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batches(rank: int, nb: int, B: int, workload: str = "edos"):
    from dostransformer_b200.synthetic import make_edos_batch, make_large_cell_batch
    if workload == "large":      # BASELINE configs[3]: 200-400 atoms per crystal, 24 neighbours (not the headline line)
        return [make_large_cell_batch(B, seed=4000 + 1000 * rank + i, T=T) for i in range(nb)]
    return [make_edos_batch(B, seed=2000 + 1000 * rank + i, T=T) for i in range(nb)]


def flops_per_crystal_fwd(n_nodes: float, n_edges: float, nmax: float) -> float:
    """SURVEY.md 8d formula (forward FLOPs per crystal), H=256, L=3, t=2."""
    H, L, t = HIDDEN, GNN_LAYERS, T_LAYERS
    Fa, Fe, Fg = 200, 41, 2
    f = 2 * n_nodes * (Fa * H + H * H) + 2 * n_edges * (Fe * H + H * H) + 2 * (Fg * H + H * H)
    f += L * 2 * n_edges * (3 * H * 2 * H + 2 * H * H) + L * 2 * n_nodes * (2 * H * 2 * H + 2 * H * H)
    f += 3 * t * 4 * T * nmax * H + 2 * t * 4 * T * T * H + 5 * t * 2 * T * (8 * H * H)
    f += 2 * T * (2 * H * H) + 2 * T * (2.5 * H * H) + 4 * T * H + 2 * (2 * H * H)
    return f


def cpu_port_throughput(sample_B: int, steps: int, warmup: int, threads: int, seed: int = 2000):
    """fwd+bwd of the CPU oracle port on `sample_B` crystals of the same generator; returns (crystals/s, ms/step)."""
    from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
    from dostransformer_b200.synthetic import make_edos_batch
    from oracle import dost_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = O.state_dict_of(DOSTransformer(GNN_LAYERS, T_LAYERS, 200, 41, 2, HIDDEN, "cpu", 0.0))
    g = make_edos_batch(sample_B, seed=seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.run_train_step(O.edos_forward, O.edos_loss, sd, g, g.y_ft)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return sample_B / med, med * 1e3, sum(times)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_sample
    val, ms, _ = cpu_port_throughput(sample, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"eDOS DOSTransformer hidden={HIDDEN} L={GNN_LAYERS} t={T_LAYERS} T={T}, random-split shape; "
                               f"bounded sample of {sample} crystals per step (CPU)", "sample_crystals": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} crystals/step, median of {args.steps} steps, oracle/dost_oracle.py "
                                   "(torch CPU restatement of the reference, pinned by tests/golden)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def time_kernel(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(iters):
        fn()
    b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def kernel_rooflines(B: int, pk, precision: str):
    """Times the dominant kernels alone, on the stream they are launched on, at the shapes the step uses.

    `roofline`: the GEMM kernel of the FFN block (65% of the model's FLOPs): the six GEMMs of one FFN layer execution
    (fc1/fc2 forward, two input-gradient and two weight-gradient GEMMs), algorithmic FLOPs = 2*M*N*K each.
    `roofline_hbm`: the CSR segmented reduction that replaces scatter_sum, and the LayerNorm that writes operand planes."""
    from dostransformer_b200 import _lib as L
    from dostransformer_b200 import ops
    dev = torch.device("cuda")
    out = {}
    M, Hh, F = B * T, HIDDEN, 4 * HIDDEN
    P = L.PRECISIONS[precision]
    shapes = [("fc1 fwd", M, F, Hh, L.KC, L.KC, 1), ("fc2 fwd", M, Hh, F, L.KC, L.KC, 1), ("fc2 dA", M, F, Hh, L.KC, L.MC, 1),
              ("fc1 dA", M, Hh, F, L.KC, L.MC, 1), ("fc2 dW", Hh, F, M, L.MC, L.MC, 0), ("fc1 dW", F, Hh, M, L.MC, L.MC, 0)]
    kname = {"fp32": "gemm_kernel<float> (fp32 FMA pipe)",
             "bf16x3": "bf::gemm_bf_kernel<3,256,pairs> (TMA-fed tcgen05 cta_group::2, 3 MMAs per product: tensor-pipe work = 3x "
                       "algorithmic)",
             "bf16": "bf::gemm_bf_kernel<1,256,pairs> (TMA-fed tcgen05 cta_group::2)"}

    def six_gemms(prec):
        Pp = L.PRECISIONS[prec]
        per_shape, tot_fl, tot_s = [], 0.0, 0.0
        for name, m, n, k, am, bm, split in shapes:
            a = torch.randn((m, k) if am == L.KC else (k, m), device=dev)
            b = torch.randn((n, k) if bm == L.KC else (k, n), device=dev)
            o = torch.empty(m, n, device=dev)
            if prec == "fp32":
                fn = lambda: ops.gemm_raw(M=m, N=n, K=k, a=[(a, None)], a_mode=am, b=b, b_mode=bm, out=o,
                                          split_k=(ops._pick_split(m, n, k, 4) if split == 0 else 1), prec=Pp)
            else:
                with ops.precision(prec):
                    ap, bp = ops.split_planes(a), ops.split_planes(b)
                sk = ops._split_for(m, n, k) if split == 0 else 1
                fn = lambda: ops.gemm_planes(M=m, N=n, K=k, a=[ap], a_mode=am, b=bp, b_mode=bm, out=o, split_k=sk, prec=Pp)
            sec = time_kernel(fn)
            fl = 2.0 * m * n * k
            per_shape.append({"gemm": name, "M": m, "N": n, "K": k, "ms": sec * 1e3, "tflops": fl / sec / 1e12})
            tot_fl += fl
            tot_s += sec
            del a, b, o
        tf = tot_fl / tot_s / 1e12
        return {"kernel": kname[prec] + ", the six GEMMs of one FFN layer execution", "bound": "tensor", "achieved": tf,
                "peak": pk["tensor"], "unit": "TFLOP/s", "frac": tf / pk["tensor"], "traffic": NCU_TRAFFIC.get(prec),
                "peak_source": pk["source"] + ", bf16 burst", "algorithmic_flops_per_launch": tot_fl / len(shapes),
                "launch_ms": tot_s / len(shapes) * 1e3, "tensor_pipe_tflops": tf * (3 if prec == "bf16x3" else 1),
                "per_shape": per_shape}

    out["roofline"] = six_gemms(precision)
    if precision == "bf16x3":
        out["roofline_bf16_mode"] = six_gemms("bf16")     # the same kernel reading only the hi planes (1 MMA per product)
    # scatter_sum as a CSR segmented reduction: [E,256] -> [N,256]
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(B, seed=2000, T=T)
    gr = ops.build_graph(g.edge_index.to(dev), g.batch.to(dev), g.system.to(dev))
    dst = torch.empty(gr.N, HIDDEN, device=dev)
    big = [torch.randn(gr.E, HIDDEN, device=dev) for _ in range(max(1, int(300e6 // (gr.E * HIDDEN * 4))))]
    it = {"i": 0}

    def seg():
        s = big[it["i"] % len(big)]
        it["i"] += 1
        ops.segment_reduce_raw(s, gr.by_dst.rowptr, gr.by_dst.perm, gr.N, out=dst)

    sec = time_kernel(seg, iters=max(10, len(big) * 3))
    nbytes = 4.0 * HIDDEN * (gr.E + gr.N) + 4.0 * gr.E + 4.0 * (gr.N + 1)
    gbs = nbytes / sec / 1e9
    out["roofline_hbm"] = {"kernel": "segment_reduce_vec_kernel<float,2> (scatter_sum [E,256]->[N,256])", "bound": "hbm",
                           "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
                           "algorithmic_bytes_per_launch": nbytes, "launch_ms": sec * 1e3, "E": gr.E, "N": gr.N,
                           "note": "inputs rotated over >300 MB so they do not sit in L2"}
    del big
    # LayerNorm of the [B*T, 256] token stream writing bf16 hi/lo operand planes: reads 4 B, writes 4 B per element
    xs = [torch.randn(M, HIDDEN, device=dev) for _ in range(3)]
    gam, bet = torch.ones(HIDDEN, device=dev), torch.zeros(HIDDEN, device=dev)
    it2 = {"i": 0}

    def ln():
        it2["i"] += 1
        with ops.precision("bf16x3"):
            ops.ln_fwd_planes(xs[it2["i"] % 3], gam, bet)

    sec = time_kernel(ln, iters=12)
    nbytes = 8.0 * M * HIDDEN + 8.0 * M
    out["roofline_hbm_ln"] = {"kernel": "rbf::ln_fwd_kernel<2> (LayerNorm [B*T,256] -> bf16 hi/lo planes)", "bound": "hbm",
                              "achieved": nbytes / sec / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                              "frac": nbytes / sec / 1e9 / pk["hbm"], "traffic": None, "algorithmic_bytes_per_launch": nbytes,
                              "launch_ms": sec * 1e3}
    return out


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of the fc2-forward
# GEMM (profiles/r1_ncu_gemm_bf_pairs_fc2_x3.summary.txt); algorithmic bytes of that launch: 421 MB planes + 105 MB output.
NCU_TRAFFIC = {"bf16x3": 507.4e6, "bf16": None, "fp32": None}


def run_product(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    from dostransformer_b200 import _lib as L
    from dostransformer_b200 import ops
    from dostransformer_b200.dp import GradReducer, live_named_parameters
    from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; dostransformer_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.lib()
    pk = peaks()
    B = args.batch
    NB = 3
    host = [b.pin_memory() for b in make_batches(rank if args.data_rank < 0 else args.data_rank, NB, B, args.workload)]
    nmax = max(int(torch.bincount(b.batch).max()) for b in host)
    n_nodes = sum(b.batch.numel() for b in host) / NB
    n_edges = sum(b.edge_index.shape[1] for b in host) / NB
    if world > 1:
        t = torch.tensor([nmax], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nmax = int(t.item())
    if args.nmax > 0:
        nmax = max(nmax, args.nmax)
    torch.manual_seed(0)
    model = DOSTransformer(GNN_LAYERS, T_LAYERS, 200, 41, 2, HIDDEN, dev, 0.0, n_energies=T, precision=args.precision).to(dev).train()
    model.max_num_nodes = nmax            # global padding length: the only cross-rank coupling besides the grads
    reducer = GradReducer(live_named_parameters(model)) if (world > 1 and not os.environ.get("DOST_BENCH_NO_REDUCER")) else None
    weight = 1.0 / world
    resident = [b.clone().to(dev) for b in host]

    def step(g):
        model.zero_grad(set_to_none=True)
        dg, _, ds = model(g)
        loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos", beta=1.0)
        if world > 1:
            loss = loss * weight
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(resident[i % NB])
    barrier()
    if rank == 0:
        sampler.wait_ready()     # a slow NVML attach must finish outside the timed region
    barrier()
    sampler.mark()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.launch_count()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record(st)
    for i in range(args.steps):
        if args.ncu_window and i == 0:       # `ncu --profile-from-start off python bench.py --ncu-window`: the launch
            torch.cuda.cudart().cudaProfilerStart()      # list of exactly one timed step of this very command
        step(resident[i % NB])
        if args.ncu_window and i == 0:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        marks[i].record(st)
    e1.record(st)
    barrier()
    dev_each = [round(a.elapsed_time(b), 2) for a, b in zip([e0] + marks[:-1], marks)]
    launches = L.launch_count() - l0
    sec = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    value = world * B * args.steps / sec

    # ---------------------------------------------------------------- end-to-end timing (host batches)
    h2d = host[0].nbytes()
    for i in range(max(NB, min(2, args.warmup))):       # every distinct batch shape once: allocator warm-up
        step(_to_device(host[i % NB], dev)).item()
    # Long-lived objects (model, batches, autograd metadata) leave the cyclic GC's working set: without this a
    # generation-2 collection lands inside the synchronous loop every few steps and stalls one step by 10-70 ms
    # (the device-resident loop above hides such pauses behind the launch queue).
    gc.collect()
    gc.freeze()
    barrier()
    per_step = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        ts = time.perf_counter()
        loss = step(_to_device(host[i % NB], dev))      # H2D of the whole batch from pinned memory
        loss.item()                                     # D2H of the step's result
        per_step.append((time.perf_counter() - ts) * 1e3)
    barrier()
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e_val = world * B * args.steps / e2e_sec

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3 operands, f32 accumulate", "bf16": "bf16 operands, f32 accumulate"}[
            args.precision], "data": "synthetic",
        "config": {"workload": (f"eDOS DOSTransformer hidden={HIDDEN} L={GNN_LAYERS} t={T_LAYERS} T={T}, random-split shape "
                                f"(BASELINE configs[1]/[2]), {B} crystals per GPU") if args.workload == "edos" else
                               (f"eDOS DOSTransformer hidden={HIDDEN} L={GNN_LAYERS} t={T_LAYERS} T={T}, large-cell stress shape "
                                f"(BASELINE configs[3]: 200-400 atoms, 24 neighbours), {B} crystals per GPU"),
                   "crystals_per_gpu": B,
                   "global_batch": B * world, "parallelism": f"dp{world}", "mean_nodes_per_batch": n_nodes,
                   "mean_edges_per_batch": n_edges, "nmax": nmax, "precision": PREC_DESC[args.precision],
                   "l2_policy": "3 distinct batches rotated; per-step activations (>1 GB) exceed the 126 MB L2"},
        "clocks": clocks, "gpu_launches": int(launches), "ms_each_step_device": dev_each,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_sec / args.steps * 1e3, "ms_each_step": [round(x, 2) for x in per_step],
                "api": "DOSTransformer(batch) + ops.dos_loss + loss.backward(), batch copied from pinned host memory, "
                       "loss.item() every step; gc.freeze() after warm-up"},
    }
    fl = 3.0 * B * flops_per_crystal_fwd(n_nodes / B, n_edges / B, nmax)
    line["model_tflops"] = fl * args.steps / sec / 1e12
    if world == 1:
        # BASELINE config 5 shape: forward-only DOS prediction, every crystal evaluated without padding (the reference's
        # eval loaders use batch_size 1), many crystals per launch, nothing read back per batch
        model.eval()
        model.per_crystal_eval = True
        with torch.no_grad():
            for i in range(2):
                model(resident[i % NB])
            torch.cuda.synchronize()
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record(st)
            for i in range(args.steps):
                model(resident[i % NB])
            i1.record(st)
            torch.cuda.synchronize()
        isec = i0.elapsed_time(i1) * 1e-3
        line["inference"] = {"value": B * args.steps / isec, "unit": UNIT, "ms_per_batch": isec / args.steps * 1e3,
                             "mode": "forward only, eval, per-crystal (no phantom keys), B crystals per launch"}
        model.per_crystal_eval = False
        model.train()
        # the step that follows the hot path (not part of the metric): fused AdamW over the live parameters
        from dostransformer_b200.optim import AdamW
        opt = AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)
        step(resident[0])
        opt.step()
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(st)
        for _ in range(10):
            opt.step()
        o1.record(st)
        torch.cuda.synchronize()
        nlive = sum(p.numel() for p in model.parameters() if p.grad is not None)
        oms = o0.elapsed_time(o1) / 10
        line["optimizer"] = {"kind": "dost_adamw_step (fused multi-tensor AdamW, lr 1e-4, wd 1e-2)", "ms_per_step": oms,
                             "live_parameters": nlive, "gbytes_per_s": 28.0 * nlive / (oms * 1e-3) / 1e9}
        torch.manual_seed(0)     # the optimizer steps above moved the weights; nothing below depends on their values
    if world == 1:
        try:
            line["device_collate"] = device_collate_bench(host, dev, B, step, args.steps, model)
        except Exception as ex:
            line["device_collate"] = {"error": repr(ex)}
    if world == 1 and args.precision == "bf16x3" and not args.no_alt:
        # the same step with plain bf16 operands (hi plane only), for reference: stated tolerance 5e-2 on gradients
        model.precision = "bf16"
        for i in range(3):
            step(resident[i % NB])
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(st)
        for i in range(args.steps):
            step(resident[i % NB])
        a1.record(st)
        torch.cuda.synchronize()
        alt = a0.elapsed_time(a1) * 1e-3
        line["bf16_mode"] = {"value": B * args.steps / alt, "unit": UNIT, "ms_per_step": alt / args.steps * 1e3,
                             "precision": PREC_DESC["bf16"]}
        model.precision = args.precision
    if world == 1:
        try:
            line.update(kernel_rooflines(B, pk, args.precision))
        except Exception as ex:  # keep the headline even if the side measurement fails
            line["roofline"] = {"error": repr(ex)}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            val, ms, total = cpu_port_throughput(args.cpu_sample, args.cpu_steps, 1, threads)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_sample} crystals/step of the same generator, median of "
                                              f"{args.cpu_steps} fwd+bwd steps after 1 warm-up ({total:.1f} s of CPU work), "
                                              "oracle/dost_oracle.py (torch CPU restatement of the reference)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def device_collate_bench(host, dev, B, step, steps, model):
    """the batches' crystals packed once into HBM-resident tables; a step sends B crystal ids and the batch is
    assembled on the device (csrc/collate.cu).  Reports the assembly alone (CUDA events, bytes = read + written) and the
    end-to-end step that starts from host ids."""
    from dostransformer_b200.collate import PackedCrystals, split_batch
    graphs = [g for b in host for g in split_batch(b)]
    pk = PackedCrystals.from_graphs(graphs, device=dev)
    gen = torch.Generator().manual_seed(7)
    ids = [torch.randperm(len(pk), generator=gen)[:B].pin_memory() for _ in range(6)]
    for i in ids[:3]:
        b = pk.collate(i)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    c0.record(st)
    for r in range(reps):
        b = pk.collate(ids[r % len(ids)])
    c1.record(st)
    torch.cuda.synchronize()
    cms = c0.elapsed_time(c1) / reps
    nbytes = 2.0 * b.nbytes()
    for i in ids:                                       # every batch shape once: allocator warm-up
        step(pk.collate(i)).item()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(steps):
        g = pk.collate(ids[r % len(ids)])
        step(g).item()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    # BASELINE configs[4] shape at store size: forward-only sweep over every crystal of the store, batches assembled on the
    # device in length-sorted order, per-crystal evaluation, predictions kept on the device
    from dostransformer_b200.evaluate import sweep
    sweep(model, pk, batch_size=B)
    torch.cuda.synchronize()
    reps_s = 4
    ts = time.perf_counter()
    for _ in range(reps_s):
        ids_s, ds_s, _ = sweep(model, pk, batch_size=B)
    torch.cuda.synchronize()
    sweep_sec = (time.perf_counter() - ts) / reps_s
    return {"value": B * steps / sec, "unit": UNIT, "ms_per_step": sec / steps * 1e3, "h2d_bytes_per_step": 8 * B,
            "inference_sweep": {"value": len(pk) / sweep_sec, "unit": UNIT, "crystals": len(pk), "batch": B,
                                "note": "evaluate.sweep: ids -> on-device collate (length-sorted batches) -> forward, "
                                        "per-crystal evaluation, host wall clock incl. launch overhead"},
            "d2h_bytes_per_step": 4, "collate_ms": cms, "collate_gbytes_per_s": nbytes / (cms * 1e-3) / 1e9,
            "collate_launches": len(pk.tables) + 2, "store_bytes": pk.nbytes(), "crystals_in_store": len(pk),
            "note": "crystal ids from host memory -> PackedCrystals.collate (segmented copies on the device) -> fwd+bwd; "
                    "collate_ms includes the host-side launch cost of its kernels"}


def _to_device(g, dev):
    from dostransformer_b200.synthetic import CrystalBatch
    return CrystalBatch(**{k: (getattr(g, k).to(dev, non_blocking=True) if torch.is_tensor(getattr(g, k)) else
                               getattr(g, k)) for k in g.keys()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="crystals per GPU")
    ap.add_argument("--cpu-sample", type=int, default=64, help="crystals per step of the CPU baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=20, help="timed CPU baseline steps (bounded sample, ~10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra bf16-mode measurement")
    ap.add_argument("--workload", default="edos", choices=["edos", "large"],
                    help="edos = the headline configuration; large = BASELINE configs[3] (use --batch 64)")
    ap.add_argument("--nmax", type=int, default=0, help="(experiments) force a larger global padding length")
    ap.add_argument("--data-rank", type=int, default=-1, help="(experiments) generate the batches of another rank")
    ap.add_argument("--precision", default=os.environ.get("DOST_PRECISION", "bf16x3"), choices=["fp32", "bf16x3", "bf16"],
                    help="GEMM path: fp32 FMA pipe | tcgen05 bf16x3 (fp32 parity, default) | tcgen05 bf16")
    ap.add_argument("--ncu-window", action="store_true", help="bracket the first timed step with cudaProfilerStart/Stop (for "
                    "ncu --profile-from-start off; a number printed under a profiler is not a bench value)")
    ap.add_argument("--energies", type=int, default=T, help="energy-grid length (201 = the reference's eDOS grid; 1001 = the "
                    "long-grid variant of the large-cell stress configuration); not the headline line when changed")
    args = ap.parse_args()
    if args.energies != T:
        globals()["T"] = args.energies
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_product(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

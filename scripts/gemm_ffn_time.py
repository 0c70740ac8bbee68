"""Times the six GEMMs of one FFN layer execution exactly as ops._FFNBlock issues them (shape AND epilogue) with CUDA
events, for each value of an environment switch read per call by the library.
usage: gemm_ffn_time.py [ENVVAR=v1,v2 ...] [--prec bf16x3] [--iters 20]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops

prec, iters, sweeps = "bf16x3", 20, []
args = sys.argv[1:]
while args:
    a = args.pop(0)
    if a == "--prec":
        prec = args.pop(0)
    elif a == "--iters":
        iters = int(args.pop(0))
    elif "=" in a:
        k, v = a.split("=")
        sweeps.append((k, v.split(",")))
B, T, H = 512, 201, 256
M, F = 2 * B * T, 4 * H
dev = "cuda"
NBUF = int(os.environ.get("NBUF", "3"))      # rotate over > L2-sized operand sets
with ops.precision(prec):
    ys = [torch.randn(M, H, device=dev) for _ in range(NBUF)]
    h0 = [ops.split_planes(torch.randn(M, H, device=dev)) for _ in range(NBUF)]
    h1 = [ops.split_planes(torch.randn(M, F, device=dev)) for _ in range(NBUF)]
    w1p, w2p = ops.split_planes(torch.randn(F, H, device=dev) * 0.05), ops.split_planes(torch.randn(H, F, device=dev) * 0.05)
    b1, b2 = torch.randn(F, device=dev), torch.randn(H, device=dev)
    dop = [ops.split_planes(torch.randn(M, H, device=dev)) for _ in range(NBUF)]
    dv1 = [ops.split_planes(torch.randn(M, F, device=dev)) for _ in range(NBUF)]
    o_mh, db1 = torch.empty(M, H, device=dev), torch.empty(F, device=dev)
    gates = [torch.zeros(F // 32, M, dtype=torch.int32, device=dev) for _ in range(NBUF)] if os.environ.get("GATES", "1") == "1" else None
    dw2, dw1 = torch.empty(H, F, device=dev), torch.empty(F, H, device=dev)
    fns = {
        "fc1_fwd": lambda i: ops.gemm_planes(M=M, N=F, K=H, a=[h0[i]], a_mode=L.KC, b=w1p, b_mode=L.KC, bias=b1, act=L.ACT_RELU, out_planes=h1[i],
                                             out_gate=gates[i] if gates else None),
        "fc2_fwd": lambda i: ops.gemm_planes(M=M, N=H, K=F, a=[h1[i]], a_mode=L.KC, b=w2p, b_mode=L.KC, bias=b2, residual=ys[i], out=o_mh),
        "fc2_dA": lambda i: ops.gemm_planes(M=M, N=F, K=H, a=[dop[i]], a_mode=L.KC, b=w2p, b_mode=L.MC, dact=None if gates else h1[i],
                                            dact_gate=gates[i] if gates else None, dact_slope=0.0, out_planes=dv1[i], colsum_out=db1),
        "fc1_dA": lambda i: ops.gemm_planes(M=M, N=H, K=F, a=[dv1[i]], a_mode=L.KC, b=w1p, b_mode=L.MC, out=o_mh),
        "fc2_dW": lambda i: ops.gemm_planes(M=H, N=F, K=M, a=[dop[i]], a_mode=L.MC, b=h1[i], b_mode=L.MC, out=dw2, split_k=ops._split_for(H, F, M)),
        "fc1_dW": lambda i: ops.gemm_planes(M=F, N=H, K=M, a=[dv1[i]], a_mode=L.MC, b=h0[i], b_mode=L.MC, out=dw1, split_k=ops._split_for(F, H, M)),
    }
    flops = 2.0 * M * F * H

    def run_all(tag):
        line = [tag]
        for name, fn in fns.items():
            for i in range(3):
                fn(i % NBUF)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                fn(i % NBUF)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            line.append(f"{name} {ms:.4f} ms {flops / ms / 1e9:.0f} TF")
        print(" | ".join(line), flush=True)

    if not sweeps:
        run_all("default")
    for k, vals in sweeps:
        for rep in range(2):
            for v in vals:
                os.environ[k] = v
                run_all(f"{k}={v}")
        os.environ.pop(k, None)

"""Where does the K=256 FFN GEMM spend its time?  Same M, N; K swept; epilogue variants.  (diagnostic, bf16x3)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops

dev = "cuda"
M, N = 512 * 201, 1024


def timeit(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ONE = (int(sys.argv[2]), sys.argv[3]) if len(sys.argv) > 3 and sys.argv[1] == "--one" else None   # --one <K> <fp32|planes|hi>
out = torch.empty(M, N, device=dev)
op = ops.empty_planes(M, N, dev, True)
oph = ops.empty_planes(M, N, dev, False)
bias = torch.randn(N, device=dev)
mask = ops.split_planes(torch.randn(M, N, device=dev))
csum = torch.empty(N, device=dev)
with ops.precision("bf16x3"):
    for K in ((ONE[0],) if ONE else (64, 128, 256, 512, 1024)):
        a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
        ap, bp = ops.split_planes(a), ops.split_planes(b)
        base = dict(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=bp, b_mode=L.KC)
        variants = {"fp32 out": dict(out=out), "planes hi+lo, bias+relu": dict(out_planes=op, bias=bias, act=L.ACT_RELU),
                    "hi plane only": dict(out_planes=oph),
                    "planes, dact+colsum (fc2 dA)": dict(out_planes=op, dact=mask, dact_slope=0.0, colsum_out=csum),
                    "planes, dact": dict(out_planes=op, dact=mask, dact_slope=0.0),
                    "planes, colsum": dict(out_planes=op, colsum_out=csum)}
        if ONE:
            key = {"fp32": "fp32 out", "planes": "planes hi+lo, bias+relu", "hi": "hi plane only"}[ONE[1]]
            for _ in range(4):
                ops.gemm_planes(**base, **variants[key])
            torch.cuda.synchronize()
            sys.exit(0)
        for name, kw in variants.items():
            ms = timeit(lambda: ops.gemm_planes(**base, **kw))
            print(f"N={N} K={K:5d} {name:30s} {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s (x3 effective)", flush=True)
    # the same output volume as 4 column tiles of N=256 (K=256): tile count unchanged, B operand 4x smaller
    K = 256
    a, b = torch.randn(M, K, device=dev), torch.randn(256, K, device=dev)
    ap, bp = ops.split_planes(a), ops.split_planes(b)
    o2 = torch.empty(M, 256, device=dev)
    ms = timeit(lambda: ops.gemm_planes(M=M, N=256, K=K, a=[ap], a_mode=L.KC, b=bp, b_mode=L.KC, out=o2))
    print(f"N=256 K=256 fp32 out {ms:7.3f} ms  {2.0 * M * 256 * K / ms / 1e9:7.1f} TFLOP/s", flush=True)

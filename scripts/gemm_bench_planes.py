"""Times dost_gemm_bf16 (TMA-fed tcgen05 over bf16 planes) at the step's dominant shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops

dev = "cuda"
B, T, H = 512, 201, 256
E = 144004
shapes = [("fc1 fwd   KC/KC", B * T, 4 * H, H, L.KC, L.KC, 1), ("fc2 fwd   KC/KC", B * T, H, 4 * H, L.KC, L.KC, 1),
          ("fc1 dA    KC/MC", B * T, H, 4 * H, L.KC, L.MC, 1), ("fc1 dW    MC/MC", 4 * H, H, B * T, L.MC, L.MC, 37),
          ("fc2 dA    KC/MC", B * T, 4 * H, H, L.KC, L.MC, 1), ("fc2 dW    MC/MC", H, 4 * H, B * T, L.MC, L.MC, 37),
          ("edge1 fwd KC/KC", E, 2 * H, 3 * H, L.KC, L.KC, 1), ("edge1 dW  MC/MC", 2 * H, 3 * H, E, L.MC, L.MC, 24)]
precs = sys.argv[1:] or ["bf16x3", "bf16"]
for name, M, N, K, am, bm, split in shapes:
    a = torch.randn((M, K) if am == L.KC else (K, M), device=dev)
    b = torch.randn((N, K) if bm == L.KC else (K, N), device=dev)
    out = torch.empty(M, N, device=dev)
    for prec in precs:
        with ops.precision(prec):
            ap, bp = ops.split_planes(a), ops.split_planes(b)
            fn = lambda: ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=am, b=bp, b_mode=bm, out=out, split_k=split)
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            st = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                fn()
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
        print(f"{name:18s} M={M:7d} N={N:5d} K={K:7d} {prec:7s} {ms:8.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)

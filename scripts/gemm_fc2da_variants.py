"""Which feature of the FFN's fc2 input-gradient GEMM costs what: times the launch with features removed one at a time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops

B, T, H = 512, 201, 256
M, F = 2 * B * T, 4 * H
dev = "cuda"
with ops.precision("bf16x3"):
    h1 = ops.split_planes(torch.randn(M, F, device=dev))
    w2p = ops.split_planes(torch.randn(H, F, device=dev) * 0.05)        # [H, F]: MC operand of dout @ W2
    w2t = ops.split_planes((torch.randn(F, H, device=dev) * 0.05))      # [F, H]: the same product with a KC operand
    dop = ops.split_planes(torch.randn(M, H, device=dev))
    dv1 = ops.empty_planes(M, F, dev, True)
    o32 = torch.empty(M, F, device=dev)
    db1 = torch.empty(F, device=dev)
    b1 = torch.randn(F, device=dev)
    variants = {
        "full (MC B, mask, planes out, colsum)": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2p, b_mode=L.MC, dact=h1, dact_slope=0.0, out_planes=dv1, colsum_out=db1),
        "no colsum": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2p, b_mode=L.MC, dact=h1, dact_slope=0.0, out_planes=dv1),
        "no mask": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2p, b_mode=L.MC, out_planes=dv1, colsum_out=db1),
        "no mask, no colsum": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2p, b_mode=L.MC, out_planes=dv1),
        "KC B, mask, colsum": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2t, b_mode=L.KC, dact=h1, dact_slope=0.0, out_planes=dv1, colsum_out=db1),
        "KC B, plain planes out (= fc1 fwd without bias/relu)": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2t, b_mode=L.KC, out_planes=dv1),
        "KC B, bias+relu, planes out (= fc1 fwd)": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2t, b_mode=L.KC, bias=b1, act=L.ACT_RELU, out_planes=dv1),
        "KC B, hi plane only out (bf16 planes target)": None,
        "KC B, fp32 out": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2t, b_mode=L.KC, out=o32),
    }
    for rep in range(2):
        for name, fn in variants.items():
            if fn is None:
                continue
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"{e0.elapsed_time(e1) / 10:.4f} ms  {name}", flush=True)

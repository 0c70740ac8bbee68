"""Runs one GEMM of the FFN block exactly as ops._FFNBlock issues it (shape AND epilogue), a few times, for ncu captures.
usage: gemm_ffn_one.py <fc1_fwd|fc2_fwd|fc2_dA|fc1_dA|fc2_dW|fc1_dW> <bf16x3|bf16> [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops

name, prec = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
B, T, H = 512, 201, 256
M, F = 2 * B * T, 4 * H
dev = "cuda"
with ops.precision(prec):
    y = torch.randn(M, H, device=dev)
    h0p = ops.split_planes(torch.randn(M, H, device=dev))
    h1p = ops.split_planes(torch.randn(M, F, device=dev))
    w1p, w2p = ops.split_planes(torch.randn(F, H, device=dev) * 0.05), ops.split_planes(torch.randn(H, F, device=dev) * 0.05)
    b1, b2 = torch.randn(F, device=dev), torch.randn(H, device=dev)
    dop = ops.split_planes(torch.randn(M, H, device=dev))
    dv1p = ops.split_planes(torch.randn(M, F, device=dev))
    o_mh, db1 = torch.empty(M, H, device=dev), torch.empty(F, device=dev)
    gate = torch.zeros(F // 32, M, dtype=torch.int32, device=dev)
    dw2, dw1 = torch.empty(H, F, device=dev), torch.empty(F, H, device=dev)
    fns = {
        "fc1_fwd": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[h0p], a_mode=L.KC, b=w1p, b_mode=L.KC, bias=b1, act=L.ACT_RELU, out_planes=h1p, out_gate=gate),
        "fc2_fwd": lambda: ops.gemm_planes(M=M, N=H, K=F, a=[h1p], a_mode=L.KC, b=w2p, b_mode=L.KC, bias=b2, residual=y, out=o_mh),
        "fc2_dA": lambda: ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2p, b_mode=L.MC, dact_gate=gate, dact_slope=0.0,
                                          out_planes=dv1p, colsum_out=db1),
        "fc1_dA": lambda: ops.gemm_planes(M=M, N=H, K=F, a=[dv1p], a_mode=L.KC, b=w1p, b_mode=L.MC, out=o_mh),
        "fc2_dW": lambda: ops.gemm_planes(M=H, N=F, K=M, a=[dop], a_mode=L.MC, b=h1p, b_mode=L.MC, out=dw2, split_k=ops._split_for(H, F, M)),
        "fc1_dW": lambda: ops.gemm_planes(M=F, N=H, K=M, a=[dv1p], a_mode=L.MC, b=h0p, b_mode=L.MC, out=dw1, split_k=ops._split_for(F, H, M)),
    }
    for _ in range(iters):
        fns[name]()
torch.cuda.synchronize()
print("ok", name, prec, M)

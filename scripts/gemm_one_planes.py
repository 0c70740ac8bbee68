"""Runs one dost_gemm_bf16 shape a few times (for ncu captures).  usage: gemm_one_planes.py <name> <prec>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops

B, T, H, E = 512, 201, 256, 144004
SHAPES = {"fc1": (B * T, 4 * H, H, L.KC, L.KC, 1), "fc2": (B * T, H, 4 * H, L.KC, L.KC, 1),
          "fc1_dA": (B * T, H, 4 * H, L.KC, L.MC, 1), "fc1_dW": (4 * H, H, B * T, L.MC, L.MC, 37),
          "edge1": (E, 2 * H, 3 * H, L.KC, L.KC, 1), "edge1_dW": (2 * H, 3 * H, E, L.MC, L.MC, 24)}
name, prec = sys.argv[1], sys.argv[2]
M, N, K, am, bm, split = SHAPES[name]
dev = "cuda"
a = torch.randn((M, K) if am == L.KC else (K, M), device=dev)
b = torch.randn((N, K) if bm == L.KC else (K, N), device=dev)
out = torch.empty(M, N, device=dev)
with ops.precision(prec):
    ap, bp = ops.split_planes(a), ops.split_planes(b)
    for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
        ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=am, b=bp, b_mode=bm, out=out, split_k=split)
torch.cuda.synchronize()
print("ok")

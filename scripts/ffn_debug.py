import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dostransformer_b200 import _lib as L, ops
DEV = "cuda"
torch.manual_seed(3)
M, H = 1000, 256
F = 4 * H
y = torch.randn(M, H, device=DEV)
ln_w = 1 + 0.1 * torch.randn(H, device=DEV); ln_b = 0.1 * torch.randn(H, device=DEV)
w1 = torch.randn(F, H, device=DEV) / 16; b1 = 0.1 * torch.randn(F, device=DEV)
w2 = torch.randn(H, F, device=DEV) / 32; b2 = 0.1 * torch.randn(H, device=DEV)
dout = torch.randn(M, H, device=DEV)
def rel(a, b, name):
    a = a.double(); b = b.double()
    e = (a - b).abs()
    print(f"{name:8s} maxerr/max = {e.max().item() / b.abs().max().item():.3e}   worst row {e.max(1).values.argmax().item()} col {e.max(0).values.argmax().item()}  rows>1e-3: {(e.max(1).values > 1e-3 * b.abs().max()).sum().item()}")
def rec(p):
    return p.hi[:, :p.cols].double() + (p.lo[:, :p.cols].double() if p.lo is not None else 0)
with ops.precision("bf16x3"):
    _, h0p, stats = ops.ln_fwd_planes(y, ln_w, ln_b)
    w1p, w2p = ops.split_planes(w1), ops.split_planes(w2)
    h1p = ops.empty_planes(M, F, DEV)
    ops.gemm_planes(M=M, N=F, K=H, a=[h0p], a_mode=L.KC, b=w1p, b_mode=L.KC, bias=b1, act=L.ACT_RELU, out_planes=h1p)
    out = torch.empty(M, H, device=DEV)
    ops.gemm_planes(M=M, N=H, K=F, a=[h1p], a_mode=L.KC, b=w2p, b_mode=L.KC, bias=b2, residual=y, out=out)
    h0 = torch.nn.functional.layer_norm(y.double(), (H,), ln_w.double(), ln_b.double(), 1e-5)
    h1 = torch.relu(h0 @ w1.double().T + b1.double())
    rel(rec(h0p), h0, "h0"); rel(rec(h1p), h1, "h1"); rel(out, y.double() + h1 @ w2.double().T + b2.double(), "out")
    dop = ops.split_planes(dout)
    rel(rec(dop), dout, "dop")
    dv1p = ops.empty_planes(M, F, DEV)
    ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2p, b_mode=L.MC, dact=h1p, dact_slope=0.0, out_planes=dv1p)
    dv1 = (dout.double() @ w2.double()) * (h1 > 0)
    rel(rec(dv1p), dv1, "dv1")
    dv1_nomask = torch.empty(M, F, device=DEV)
    ops.gemm_planes(M=M, N=F, K=H, a=[dop], a_mode=L.KC, b=w2p, b_mode=L.MC, out=dv1_nomask)
    rel(dv1_nomask, dout.double() @ w2.double(), "dv1raw")
    dh0 = torch.empty(M, H, device=DEV)
    ops.gemm_planes(M=M, N=H, K=F, a=[dv1p], a_mode=L.KC, b=w1p, b_mode=L.MC, out=dh0)
    rel(dh0, dv1 @ w1.double(), "dh0")
    dw1 = torch.empty(F, H, device=DEV)
    ops.gemm_planes(M=F, N=H, K=M, a=[dv1p], a_mode=L.MC, b=h0p, b_mode=L.MC, out=dw1, split_k=ops._split_for(F, H, M))
    rel(dw1, dv1.T @ h0, "dw1")
    dw2 = torch.empty(H, F, device=DEV)
    ops.gemm_planes(M=H, N=F, K=M, a=[dop], a_mode=L.MC, b=h1p, b_mode=L.MC, out=dw2, split_k=ops._split_for(H, F, M))
    rel(dw2, dout.double().T @ h1, "dw2")
    rel(ops.colsum_planes(dv1p)[None], dv1.sum(0)[None], "db1")

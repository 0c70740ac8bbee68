"""Phase timeline of CTA 0 of the fused attention kernel (needs a library built with -DDOST_ATTN_TIMELINE)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops

S, T, H = 1024, 201, 256
Lk = int(sys.argv[1]) if len(sys.argv) > 1 else 201
store = len(sys.argv) > 2
dev = "cuda"
with ops.precision("bf16x3"):
    q = torch.randn(S, T, H, device=dev)
    k = torch.randn(S, Lk, H, device=dev)
    r = torch.randn(S, T, H, device=dev)
    qp, kp = ops.split_planes(q.view(S * T, H)), ops.split_planes(k.view(S * Lk, H))
    out = torch.empty(S * T, H, device=dev)
    pp = ops.empty_planes(S * T, Lk, dev) if store else None
    for _ in range(3):
        ops._fused_attention_fwd(qp, kp, S * Lk, S, T, Lk, H, None, None, None, Lk, r.view(S * T, H), T * H, out, pp)
    torch.cuda.synchronize()
lib = L.lib()
n = 64 * 16
buf = (C.c_ulonglong * n)()
lib.dost_attn_fused_timeline.restype = C.c_int
assert lib.dost_attn_fused_timeline(buf, n) == 0
names = ["prod:Q issue", "mma:Q full", "mma:QK issued", "mma:P full", "mma:PV issued", "soft:S full", "soft:max done", "soft:sum done",
         "soft:P written", "soft:O full", "soft:epilogue done"]
t0 = buf[0]
print("Lk", Lk, "store_p", store)
for local in range(0, 8):
    row = [buf[local * 16 + i] for i in range(11)]
    print(local, " ".join(f"{names[i].split(':')[1]}={(row[i] - t0) / 1000:.2f}" for i in range(11)))

"""fp32-path gradient error of one-crystal batches against the fp64 oracle (diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from dostransformer_b200 import ops
from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
from dostransformer_b200.synthetic import make_edos_batch
from oracle import dost_oracle as O

DEV = "cuda"


def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


for B, H, seed in [(1, 128, 5), (1, 128, 7), (1, 128, 8), (1, 128, 9), (2, 128, 5), (1, 64, 5)]:
    for prec in ("fp32", "bf16x3"):
        torch.manual_seed(seed)
        m = DOSTransformer(3, 2, 200, 41, 2, H, torch.device(DEV), 0.0, precision=prec)
        sd = O.state_dict_of(m)
        g = make_edos_batch(B, seed=100 + seed, mean_atoms=10.0, max_atoms=60)
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        g64 = g.clone()
        for k in g64.keys():
            v = getattr(g64, k)
            if torch.is_tensor(v) and v.is_floating_point():
                setattr(g64, k, v.double())
        (rdg, rx, rds), rloss, rg = O.run_train_step(O.edos_forward, O.edos_loss, sd64, g64, g64.y_ft)
        _, _, rg32 = O.run_train_step(O.edos_forward, O.edos_loss, sd, g, g.y_ft)
        m.to(DEV).train()
        gd = g.clone().to(DEV)
        dg, x, ds = m(gd)
        loss = ops.dos_loss(dg, ds, gd.y_ft, mode="edos", beta=1.0)
        loss.backward()
        errs = sorted(((l2(p.grad, rg[k]), l2(rg32[k], rg[k]), k) for k, p in m.named_parameters() if p.grad is not None), reverse=True)
        print(f"B={B} H={H} seed={seed} {prec:6s} nodes={g.x.shape[0]} out: dg={l2(dg, rdg):.1e} ds={l2(ds, rds):.1e} x={l2(x, rx):.1e} "
              f"loss={abs(loss.item() - rloss.item()) / abs(rloss.item()):.1e} | worst grads: "
              + ", ".join(f"{k.split('.')[0]}..{'.'.join(k.split('.')[-2:])}={e:.1e}(cpu32 {n:.0e})" for e, n, k in errs[:4]), flush=True)

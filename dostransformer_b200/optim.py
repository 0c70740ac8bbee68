"""Fused AdamW over the model's live parameters.

Mirrors ``torch.optim.AdamW(model.parameters(), lr=args.lr, weight_decay=1e-2)`` as the reference launchers construct
it (main_eDOS.py:93, main_phDOS.py:90): same hyper-parameter names and defaults, ``step()``, ``zero_grad()``,
``state_dict()`` / ``load_state_dict()`` in torch's layout, parameters whose ``grad is None`` are skipped (so the
reference's dead parameters stay at their initial values).  One kernel launch per 32 tensors instead of torch's foreach
chain; fp32 CUDA parameters only (the phonon fp64 model keeps torch.optim.AdamW).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

import torch

from . import _lib as L


class AdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(params=self.params, **self.defaults)]
        self.state = {}
        self._step = 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def _state_of(self, p):
        st = self.state.get(p)
        if st is None:
            st = dict(step=0, exp_avg=torch.zeros_like(p, memory_format=torch.preserve_format),
                      exp_avg_sq=torch.zeros_like(p, memory_format=torch.preserve_format))
            self.state[p] = st
        return st

    @torch.no_grad()
    def step(self) -> None:
        grp = self.param_groups[0]
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        self._step += 1
        ps, gs, ms, vs, ns = [], [], [], [], []
        keep = []      # contiguous gradient copies must outlive the launch
        for p in live:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("dostransformer_b200.optim.AdamW handles contiguous fp32 CUDA parameters only")
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            keep.append(g)
            st = self._state_of(p)
            st["step"] = self._step
            ps.append(p.data_ptr()); gs.append(g.data_ptr()); ms.append(st["exp_avg"].data_ptr())
            vs.append(st["exp_avg_sq"].data_ptr()); ns.append(p.numel())
        n = len(live)
        arr = lambda xs: (C.c_void_p * n)(*xs)
        L.check(L.lib().dost_adamw_step(n, arr(ps), arr(gs), arr(ms), arr(vs), (C.c_longlong * n)(*ns), float(grp["lr"]),
                                        float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]),
                                        float(grp["weight_decay"]), self._step, L.stream()), "adamw_step")
        # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on the tensors'
        # version counters, e.g. the bf16 operand planes of the weights in ops.weight_planes) that they changed
        torch.autograd.graph.increment_version(live)

    def state_dict(self):
        idx = {p: i for i, p in enumerate(self.params)}
        return {"state": {idx[p]: {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"],
                                   "exp_avg_sq": st["exp_avg_sq"]} for p, st in self.state.items()},
                "param_groups": [{**{k: v for k, v in self.param_groups[0].items() if k != "params"},
                                  "params": list(range(len(self.params)))}]}

    def load_state_dict(self, sd) -> None:
        for i, st in sd["state"].items():
            p = self.params[int(i)]
            self.state[p] = dict(step=int(float(st["step"])), exp_avg=st["exp_avg"].to(p.device).clone(),
                                 exp_avg_sq=st["exp_avg_sq"].to(p.device).clone())
            self._step = max(self._step, self.state[p]["step"])
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v

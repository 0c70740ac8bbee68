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
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        params = list(params)
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        # torch.optim accepts either parameters or param-group dicts ({"params": [...], "lr": ...})
        groups = params if isinstance(params[0], dict) else [dict(params=params)]
        self.param_groups = []
        seen = set()
        for g in groups:
            ps = [g["params"]] if isinstance(g["params"], torch.Tensor) else list(g["params"])
            for p in ps:
                if id(p) in seen:
                    raise ValueError("some parameters appear in more than one parameter group")
                seen.add(id(p))
            extra = set(g) - {"params"} - set(self.defaults)
            if extra:
                raise ValueError(f"unknown param-group options {sorted(extra)}")
            self.param_groups.append({**self.defaults, **{k: v for k, v in g.items() if k != "params"}, "params": ps})
        self.state = {}

    @property
    def params(self) -> List[torch.nn.Parameter]:
        return [p for g in self.param_groups for p in g["params"]]

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
        """One AdamW update of every parameter that has a gradient.  The step count (bias correction) is kept PER
        PARAMETER like torch.optim.AdamW: a parameter whose gradient was None for some steps is corrected with its own
        count.  Parameters of one group that share a count go into the same launches (32 tensors per launch)."""
        touched = []
        for grp in self.param_groups:
            by_step = {}
            for p in grp["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("dostransformer_b200.optim.AdamW handles contiguous fp32 CUDA parameters only")
                st = self._state_of(p)
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            for step, live in by_step.items():
                ps, gs, ms, vs, ns = [], [], [], [], []
                keep = []      # contiguous gradient copies must outlive the launch
                for p in live:
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    keep.append(g)
                    st = self.state[p]
                    ps.append(p.data_ptr()); gs.append(g.data_ptr()); ms.append(st["exp_avg"].data_ptr())
                    vs.append(st["exp_avg_sq"].data_ptr()); ns.append(p.numel())
                n = len(live)
                arr = lambda xs: (C.c_void_p * n)(*xs)
                with torch.cuda.device(live[0].device):
                    L.check(L.lib().dost_adamw_step(n, arr(ps), arr(gs), arr(ms), arr(vs), (C.c_longlong * n)(*ns),
                                                    float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1]),
                                                    float(grp["eps"]), float(grp["weight_decay"]), step, L.stream()),
                            "adamw_step")
                touched += live
        if touched:
            # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on the tensors'
            # version counters, e.g. the bf16 operand planes of the weights in ops.weight_planes) that they changed
            torch.autograd.graph.increment_version(touched)

    @torch.no_grad()
    def step_subset(self, params: List[torch.nn.Parameter], grads: List[torch.Tensor]) -> None:
        """AdamW update of ``params`` with explicitly given (contiguous fp32) gradient tensors, on the CURRENT stream.
        Used by dp.GradReducer.finish(optimizer=...): each bucket is updated straight from its all-reduced flat buffer as
        soon as its reduction completes, so the update of bucket i overlaps the reduction of bucket i+1 and the
        write-back copy into ``p.grad`` disappears (AdamW fused with the all-reduce epilogue)."""
        group_of = {id(p): g for g in self.param_groups for p in g["params"]}
        by_key = {}
        for p, gr in zip(params, grads):
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and gr.is_contiguous()):
                raise RuntimeError("dostransformer_b200.optim.AdamW handles contiguous fp32 CUDA parameters only")
            st = self._state_of(p)
            st["step"] += 1
            by_key.setdefault((id(group_of[id(p)]), st["step"]), []).append((p, gr))
        for (gid, step), items in by_key.items():
            grp = group_of[id(items[0][0])]
            n = len(items)
            arr = lambda xs: (C.c_void_p * n)(*xs)
            L.check(L.lib().dost_adamw_step(
                n, arr([p.data_ptr() for p, _ in items]), arr([g.data_ptr() for _, g in items]),
                arr([self.state[p]["exp_avg"].data_ptr() for p, _ in items]),
                arr([self.state[p]["exp_avg_sq"].data_ptr() for p, _ in items]),
                (C.c_longlong * n)(*[p.numel() for p, _ in items]), float(grp["lr"]), float(grp["betas"][0]),
                float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]), step, L.stream()), "adamw_step")
        torch.autograd.graph.increment_version(list(params))

    def state_dict(self):
        plist = self.params
        idx = {p: i for i, p in enumerate(plist)}
        groups, k = [], 0
        for g in self.param_groups:
            n = len(g["params"])
            groups.append({**{kk: v for kk, v in g.items() if kk != "params"}, "params": list(range(k, k + n))})
            k += n
        return {"state": {idx[p]: {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"],
                                   "exp_avg_sq": st["exp_avg_sq"]} for p, st in self.state.items()},
                "param_groups": groups}

    def load_state_dict(self, sd) -> None:
        plist = self.params
        for i, st in sd["state"].items():
            p = plist[int(i)]
            self.state[p] = dict(step=int(float(st["step"])), exp_avg=st["exp_avg"].to(p.device).clone(),
                                 exp_avg_sq=st["exp_avg_sq"].to(p.device).clone())
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            for k, v in saved.items():
                if k != "params":
                    g[k] = v

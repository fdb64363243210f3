"""Flat parameter / gradient storage and the fused Adam step for the adapter + router parameters (SURVEY.md §8 rows e, f3).

``FlatParameters`` re-homes every trainable tensor of a module as a view of ONE contiguous fp32 buffer and owns a second flat
buffer for the gradients.  After backward ``gather_grads()`` moves the per-tensor gradients into it with one multi-tensor copy
(autograd hands freshly produced gradient tensors to ``.grad`` for free; pre-installed ``.grad`` views would cost an extra
accumulate kernel per tensor), then

  * the data-parallel gradient all-reduce runs on the flat buffer in place (avmoe_b200/dist.py) and nothing is ever unpacked:
  * ``FusedAdam.step()`` consumes the reduced buffer directly -- a single streaming CUDA pass over (param, grad, exp_avg,
    exp_avg_sq) per learning-rate group instead of ~1,000 per-tensor launch groups, with the 1/world of the gradient average
    folded in -- and the parameters it updates ARE the module's parameters (views).

Semantics are torch.optim.Adam's (the reference's optimizer, AVMOE/AVE/main_trans_v3.py:311-323: ``optim.Adam(param_group)`` with
per-parameter ``lr`` groups and a StepLR scheduler on top); ``param_groups`` exposes ``lr`` so ``torch.optim.lr_scheduler`` works.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch


class FlatParameters:
    """Views of one flat fp32 buffer for parameters (``flat``) and one for gradients (``grad``), in registration order.
    Tensors are padded to 4 elements so every view (and every group range) starts 16-byte aligned."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        assert all(p.dtype == torch.float32 and p.device == dev for p in self.params), "flat storage holds fp32 tensors of one device"
        self.offsets: List[int] = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
        self._grad_views = [self.grad[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]

    def gather_grads(self, clear: bool = True) -> torch.Tensor:
        """per-tensor ``.grad`` -> the flat gradient buffer (one multi-tensor copy); parameters without a gradient contribute zeros.
        clear=True drops the per-tensor gradients afterwards (the next backward installs fresh ones at no cost)."""
        have = [(v, p.grad) for v, p in zip(self._grad_views, self.params) if p.grad is not None]
        if len(have) < len(self.params):
            self.grad.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if clear:
            for p in self.params:
                p.grad = None
        return self.grad

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def check_views(self) -> bool:
        return all(p.data_ptr() == self.flat.data_ptr() + 4 * off for p, off in zip(self.params, self.offsets))

    def range_of(self, params: Sequence[torch.nn.Parameter]) -> Tuple[int, int]:
        """[start, end) of a run of parameters that are consecutive in the flat order"""
        pos = {id(p): i for i, p in enumerate(self.params)}
        idx = sorted(p if isinstance(p, int) else pos[id(p)] for p in params)
        assert idx == list(range(idx[0], idx[-1] + 1)), "a learning-rate group must be a consecutive run of parameters"
        end = self.offsets[idx[-1]] + (self.params[idx[-1]].numel() + 3) // 4 * 4
        return self.offsets[idx[0]], end


class FusedAdam:
    """Adam over FlatParameters: one kernel launch per learning-rate group per step."""

    def __init__(self, flat: FlatParameters, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 groups: Optional[List[Dict]] = None):
        self.flat = flat
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        if groups is None:
            groups = [{"params": list(flat.params)}]
        self.param_groups: List[Dict] = []
        for g in groups:
            d = dict(self.defaults)
            d.update(g)
            d["range"] = flat.range_of(d["params"])
            d.setdefault("initial_lr", d["lr"])
            self.param_groups.append(d)
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.step_count = 0

    def zero_grad(self, set_to_none: bool = False):
        self.flat.zero_grad()

    def step(self, grad_scale: float = 1.0, zero_grad: bool = False):
        from . import _lib, ops
        if not self.flat.flat.is_cuda:
            raise RuntimeError("avmoe_b200.FusedAdam has no CPU path: parameters must live on a CUDA device")
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device_of(self.flat.flat):
            for g in self.param_groups:
                s, e = g["range"]
                b1, b2 = g["betas"]
                ptr = lambda t: C.c_void_p(t.data_ptr() + 4 * s)
                rc = lib.avmoe_adam_flat(ptr(self.flat.flat), ptr(self.flat.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), C.c_int64(e - s),
                                         C.c_float(g["lr"]), C.c_float(b1), C.c_float(b2), C.c_float(g["eps"]), C.c_float(g["weight_decay"]),
                                         C.c_float(grad_scale), C.c_int64(self.step_count), int(zero_grad), ops._stream())
                _lib.check(rc, "adam_flat")
                ops.LAUNCHES += 1

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in s.items() if k != "range"})

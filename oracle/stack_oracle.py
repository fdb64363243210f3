"""TEST INFRASTRUCTURE / CPU BASELINE — the adapter stack of a variant on the CPU oracle.

Mirrors ``avmoe_b200.stack.AdapterStack`` (same site list, same residual chain, same loss) on top of
``oracle.avmoe_oracle`` (or, where /root/reference exists and ``use_reference=True``, on the UNMODIFIED reference
modules).  Used only by bench.py's cpu_baseline / ``--impl reference`` leg and by tests.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Tuple

import torch

from . import avmoe_oracle as O

STAGE_SHAPES = O.STAGE_SHAPES


def tokens_per_segment(variant: str) -> int:
    tot = 0
    for stage, blocks in O.VARIANT_SITES[variant]:
        _, Nv, _, Na = STAGE_SHAPES[stage]
        tot += blocks * 2 * (Nv + Na)
    return tot


class OracleStack:
    def __init__(self, variant: str = "ave", seed: int = 0, use_reference: bool = False, device=None):
        self.variant = variant
        self.use_reference = use_reference
        self.device = device
        self.sites: List[Tuple[int, int]] = []
        self.adapters = []   # per site: dict(name -> (spec, state_dict or module))
        k = 0
        for stage, blocks in O.VARIANT_SITES[variant]:
            for _ in range(blocks):
                entry = {}
                for name, modality in (("a1", "audio"), ("v1", "visual"), ("a2", "audio"), ("v2", "visual")):
                    spec = O.spec_for(variant, stage, modality)
                    sd = O.make_state_dict(spec, seed=seed * 1000 + k)
                    k += 1
                    if use_reference:
                        from . import ref_harness as H
                        mod = H.build_reference(spec, sd)
                        for m in mod.modules():
                            if isinstance(m, torch.nn.MultiheadAttention):
                                m.dropout = 0.0
                        H.make_bn_grads_contiguous(mod)
                        if device is not None:
                            mod = mod.to(device)
                        entry[name] = (spec, mod)
                    else:
                        if device is not None:
                            sd = {kk: v.to(device) for kk, v in sd.items()}
                        sd = {kk: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in kk else v.clone())
                              for kk, v in sd.items()}
                        entry[name] = (spec, sd)
                self.sites.append((stage, len(self.adapters)))
                self.adapters.append(entry)

    def _call(self, entry, x, o):
        spec, obj = entry
        if self.use_reference:
            from . import ref_harness as H
            r = H.run_reference(obj, spec, x, o, training=True)
            return r["out"], r["lb_loss"]
        nb = {}
        r = O.moe_adapter_forward(x, o, obj, spec, training=True, new_buffers=nb)
        for kk, v in nb.items():
            obj[kk] = v
        return r["out"], r["lb_loss"]

    def zero_grad(self):
        for entry in self.adapters:
            for spec, obj in entry.values():
                if self.use_reference:
                    obj.zero_grad(set_to_none=True)
                else:
                    for v in obj.values():
                        if v.is_floating_point() and v.grad is not None:
                            v.grad = None

    def step(self, tokens: Dict[int, Tuple[torch.Tensor, torch.Tensor]]):
        cur = dict(tokens)
        lb_sum = None
        for stage, i in self.sites:
            f_a, f_v = cur[stage]
            for na, nv in (("a1", "v1"), ("a2", "v2")):
                ra, la = self._call(self.adapters[i][na], f_a, f_v)
                rv, lv = self._call(self.adapters[i][nv], f_v, f_a)
                f_v = f_v + rv
                f_a = f_a + ra
                for l in (la, lv):
                    if l is not None:
                        lb_sum = l if lb_sum is None else lb_sum + l
            cur[stage] = (f_a, f_v)
        loss = None
        for stage in cur:
            f_a, f_v = cur[stage]
            term = f_a.sum() + f_v.sum()
            loss = term if loss is None else loss + term
        if lb_sum is not None:
            loss = loss + lb_sum
        loss.backward()
        return float(loss.detach())


def make_stage_tokens(variant: str, BT: int, seed: int = 1234, device=None):
    out = {}
    for stage, _ in O.VARIANT_SITES[variant]:
        Cv, Nv, Ca, Na = STAGE_SHAPES[stage]
        g = torch.Generator().manual_seed(seed + 97 * stage)
        f_a = torch.randn((BT, Na, Ca), generator=g) + torch.randn((BT, 1, Ca), generator=g)
        f_v = torch.randn((BT, Nv, Cv), generator=g) + torch.randn((BT, 1, Cv), generator=g)
        if device is not None:
            f_a, f_v = f_a.to(device), f_v.to(device)
        out[stage] = (f_a.requires_grad_(), f_v.requires_grad_())
    return out


def time_cpu(variant: str, BT: int, steps: int, warmup: int, use_reference: bool = False, threads: int = 0):
    """Returns (tokens_per_second, seconds_per_step list, cores used)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    stack = OracleStack(variant, use_reference=use_reference)
    toks = make_stage_tokens(variant, BT)
    times = []
    for it in range(warmup + steps):
        stack.zero_grad()
        for s in toks:
            for t in toks[s]:
                t.grad = None
        t0 = time.perf_counter()
        stack.step(toks)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = sorted(times)[len(times) // 2]
    return tokens_per_segment(variant) * BT / med, times, torch.get_num_threads()


def time_gpu_eager(variant: str, BT: int, steps: int, warmup: int, use_reference: bool = False, autocast_dtype=None, device="cuda"):
    """COMPARATOR, not a product path: the reference's (or the oracle port's) eager PyTorch modules on the GPU — what a
    reference user gets from ``model.cuda()`` ("the thing to beat on the same box").  fp32 parameters and
    tokens; ``autocast_dtype=torch.bfloat16`` wraps the forward in torch.autocast.  Returns (tokens/s, ms per step)."""
    stack = OracleStack(variant, use_reference=use_reference, device=device)
    toks = make_stage_tokens(variant, BT, device=device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def one():
        stack.zero_grad()
        for s in toks:
            for t in toks[s]:
                t.grad = None
        if autocast_dtype is not None:
            with torch.autocast("cuda", dtype=autocast_dtype):
                stack.step(toks)
        else:
            stack.step(toks)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        one()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    return tokens_per_segment(variant) * BT / (ms * 1e-3), ms

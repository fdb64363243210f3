"""Batch-sharded data parallelism for the adapter path: one process per GPU, adapters replicated, and ONE collective —
the all-reduce (average) of the adapter + router gradients per optimizer step.

The reference itself only has single-process ``nn.DataParallel`` (AVMOE/AVVP/main.py:420-421), whose gradient
reduce-add onto GPU 0 this replaces.  Gradients are packed into a few large flat fp32 buckets (reverse execution
order, so the first buckets are complete while the backward of earlier sites is still running) and reduced with
NCCL over NVLink; with gradient accumulation call ``reduce()`` only on the boundary micro-step
(AVMOE/AVE/main_trans_v3.py:133-138).  Works with the gloo backend on CPU for the host-logic tests."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 128 << 20, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # buckets in REVERSE registration order == the order gradients become ready in backward
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * 4
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)

    def num_bytes(self) -> int:
        return sum(p.numel() * 4 for p in self.params)

    def reduce(self, async_op: bool = False):
        """Average .grad of every parameter across ranks (in place).  Returns the work handles when async_op."""
        if self.world == 1:
            return []
        works = []
        for bi, bucket in enumerate(self.buckets):
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            n = sum(g.numel() for g in grads)
            flat = self._flat[bi]
            if flat is None or flat.numel() != n or flat.device != grads[0].device:
                flat = torch.empty(n, device=grads[0].device, dtype=torch.float32)
                self._flat[bi] = flat
            off = 0
            views = []
            for g in grads:
                views.append(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            torch._foreach_copy_(views, grads)          # multi-tensor pack: a handful of launches for ~1e3 small gradients
            w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            works.append((w, bucket, views))
        if async_op:
            return works
        self.finish(works)
        return []

    def finish(self, works):
        inv = 1.0 / self.world
        for (w, bucket, views), flat in zip(works, self._flat):
            w.wait()
            flat.mul_(inv)
            for p in bucket:
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in bucket], views)


def shard_batch(total: int, rank: int, world: int):
    """[start, end) of this rank's contiguous slice of the BT axis (even split, remainder to the first ranks)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGradientReducer:
    """All-reduce (sum) of a FlatParameters gradient buffer in place, in a few large buckets issued back to back on NCCL's stream;
    the division by the world size is NOT applied here -- FusedAdam.step(grad_scale=1/world) folds it into the update, so the
    gradient bytes cross HBM once for the collective and once for the optimizer and are never packed twice or unpacked."""

    def __init__(self, flat, bucket_bytes: int = 128 << 20, group=None):
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = flat.numel
        per = max(1, bucket_bytes // 4)
        self.ranges = [(s, min(n, s + per)) for s in range(0, n, per)]

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world

    def reduce(self, gather: bool = True):
        """gather per-tensor gradients into the flat buffer (unless the caller already did) and all-reduce it in place"""
        if gather:
            self.flat.gather_grads()
        if self.world == 1:
            return []
        works = [dist.all_reduce(self.flat.grad[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for s, e in self.ranges]
        for w in works:
            w.wait()
        return works

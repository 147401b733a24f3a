"""Data-parallel plumbing for the Video-Swin hot path (clips shard by rank, full weight
replica, ONE exchange step -- the gradient all-reduce).  Pure torch.distributed (NCCL on GPUs, gloo in the CPU
tests); there is no data-path collective.  bench.py uses DistributedDataParallel (bucketed all-reduce overlapped
with backward, as the reference does, agent.py:200-201); the helpers here are the explicit, testable form of the
same contract."""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def clip_shard(n_clips: int, rank: int, world: int) -> slice:
    """contiguous, balanced shard of a global batch of clips (first n_clips % world ranks get one extra)"""
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def _buckets(params: List[torch.Tensor], bucket_bytes: int) -> List[List[torch.Tensor]]:
    out, cur, size = [], [], 0
    for p in reversed(params):  # gradients become ready last-layer first
        if p.grad is None:
            continue
        if cur and (size + p.grad.numel() * p.grad.element_size() > bucket_bytes or p.grad.dtype != cur[0].grad.dtype):
            out.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += p.grad.numel() * p.grad.element_size()
    if cur:
        out.append(cur)
    return out


def all_reduce_gradients(params: Iterable[torch.Tensor], world_size: int | None = None,
                         bucket_bytes: int = 50 << 20) -> int:
    """average ``.grad`` of every parameter over the process group with bucketed flat all-reduces
    (all buckets are launched asynchronously, then waited).  Returns the number of collectives issued."""
    if not dist.is_initialized():
        return 0
    world = world_size or dist.get_world_size()
    if world == 1:
        return 0
    params = list(params)
    work = []
    for bucket in _buckets(params, bucket_bytes):
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        work.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket))
    for handle, flat, bucket in work:
        handle.wait()
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    return len(work)


def all_reduce_gradients_coalesced(params: Iterable[torch.Tensor]) -> int:
    """Copy-free form of the exchange step: every ``.grad`` tensor is averaged IN PLACE by one coalesced NCCL group call
    (``ReduceOp.AVG``; no flat buffer, no per-parameter copy or divide kernels).  Our autograd Functions hand each weight
    gradient to autograd as a fresh tensor, so DDP's bucket views would cost one copy kernel per parameter per step.
    Returns the number of tensors reduced.  (gloo has no AVG: it takes the SUM + divide route.)"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    if dist.get_backend() != "nccl":
        for g in grads:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            g.div_(dist.get_world_size())
        return len(grads)
    with dist._coalescing_manager(device=grads[0].device, async_ops=True) as cm:
        for g in grads:
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
    cm.wait()
    return len(grads)


class OverlappedGradReducer:
    """The exchange step overlapped with backward, still copy-free: parameters are grouped (one group per direct child of
    every ``nn.ModuleList`` / top-level sub-module, i.e. one per Swin block), and as soon as autograd has accumulated the
    gradient of the last parameter of a group the group's ``.grad`` tensors are averaged in place by ONE coalesced, asynchronous
    NCCL call.  ``finish()`` (after ``loss.backward()``) reduces whatever is left and waits for all outstanding calls."""

    def __init__(self, module: torch.nn.Module):
        self.groups: List[List[torch.nn.Parameter]] = []
        seen = set()

        def add(mod):
            ps = [p for p in mod.parameters() if p.requires_grad and id(p) not in seen]
            if ps:
                seen.update(id(p) for p in ps)
                self.groups.append(ps)

        def walk(mod):
            for child in mod.children():
                if isinstance(child, torch.nn.ModuleList) or any(isinstance(c, torch.nn.ModuleList) for c in child.children()):
                    walk(child)
                else:
                    add(child)
            add(mod)   # parameters held directly by `mod`
        walk(module)
        self._left = [len(g) for g in self.groups]
        self._launched = [False] * len(self.groups)
        self._works = []
        self._hooks = []
        for gi, g in enumerate(self.groups):
            for p in g:
                self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, gi=gi: self._ready(gi)))

    def _launch(self, gi: int) -> None:
        grads = [p.grad for p in self.groups[gi] if p.grad is not None]
        self._launched[gi] = True
        if not grads or not dist.is_initialized() or dist.get_world_size() == 1:
            return
        if dist.get_backend() != "nccl":
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM)
                g.div_(dist.get_world_size())
            return
        with dist._coalescing_manager(device=grads[0].device, async_ops=True) as cm:
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.AVG)
        self._works.append(cm)

    def _ready(self, gi: int) -> None:
        self._left[gi] -= 1
        if self._left[gi] == 0 and not self._launched[gi]:
            self._launch(gi)

    def finish(self) -> int:
        """call after backward: reduce groups whose parameters did not all receive a gradient, wait, re-arm; returns #groups"""
        for gi in range(len(self.groups)):
            if not self._launched[gi]:
                self._launch(gi)
        for cm in self._works:
            cm.wait()
        self._works.clear()
        self._left = [len(g) for g in self.groups]
        self._launched = [False] * len(self.groups)
        return len(self.groups)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks.clear()

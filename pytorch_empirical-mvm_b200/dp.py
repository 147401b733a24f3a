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
    # (torch.distributed._coalescing_manager is a private API -- present in torch 2.1 ... 2.11; without it the same reductions
    # are issued one by one.  dp.FlatGradReducer, the default exchange, uses public APIs only.)
    cmgr = getattr(dist, "_coalescing_manager", None)
    if cmgr is None:
        works = [dist.all_reduce(g, op=dist.ReduceOp.AVG, async_op=True) for g in grads]
        for w in works:
            w.wait()
        return len(grads)
    with cmgr(device=grads[0].device, async_ops=True) as cm:
        for g in grads:
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
    cm.wait()
    return len(grads)


class OverlappedGradReducer:
    """(Earlier form, kept for A/B against FlatGradReducer; it builds each collective's tensor list from the gradients present
    on the local rank, so every rank must produce gradients for the same parameters.)
    The exchange step overlapped with backward, still copy-free: parameters are grouped (one group per direct child of
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


class FlatGradReducer:
    """The exchange step as a few large messages overlapped with backward (the reference's DDP, agent.py:195-201).

    Every parameter gets a fixed slot in ONE flat buffer per dtype, laid out in reverse registration order (the order in which
    backward produces gradients).  While this object is installed, the wgrad / LayerNorm-backward kernels write each gradient
    straight into its slot (``functional.GRAD_SINK``) and autograd adopts that view as ``p.grad``; a gradient that arrives any
    other way is copied into its slot by the hook.  The buffer is cut into ``n_chunks`` contiguous, equally sized ranges; as soon
    as every parameter of a range has its gradient, the range is averaged in place by one asynchronous all-reduce.  The tensor
    list of every collective is a fixed slice of the flat buffer: it does not depend on which parameters received a gradient on
    which rank (slots without a gradient hold zeros).

        red = FlatGradReducer(model)
        for batch in loader:
            red.zero_grad()            # instead of setting p.grad = None
            loss(model(batch)).backward()
            red.finish()               # waits; p.grad are views of the averaged flat buffer
    """

    ALIGN = 64   # elements: every slot starts 128-byte aligned for 16-bit types (the kernels need 16 bytes)

    def __init__(self, module: torch.nn.Module, n_chunks: int = 4, install_sink: bool = True, unused: str = "zero"):
        """unused: what a parameter without a local gradient gets after finish() when the job is distributed.
        "zero" (default): its averaged slot (zeros if no rank had a gradient) -- every rank takes the same optimizer step and
        nothing synchronises the host; "none": stays None unless some rank had a gradient, decided through a small extra
        all-reduce of a presence bitmap that the host must read (one device synchronisation per step, like DDP's
        find_unused_parameters)."""
        assert unused in ("zero", "none")
        self.unused = unused
        self.params = [p for p in reversed(list(module.parameters())) if p.requires_grad]
        self.flat = {}          # dtype -> flat tensor
        self.slot = {}          # id(p) -> (dtype, offset, numel)
        self._by_ptr = {}       # data_ptr -> parameter
        sizes = {}
        for p in self.params:
            off = sizes.get(p.dtype, 0)
            self.slot[id(p)] = (p.dtype, off, p.numel())
            sizes[p.dtype] = off + (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self._by_ptr[p.data_ptr()] = p
        dev = self.params[0].device if self.params else torch.device("cpu")
        for dt, n in sizes.items():
            self.flat[dt] = torch.zeros(n, dtype=dt, device=dev)
        # chunks: contiguous ranges of each flat buffer with about equal bytes, cut at slot boundaries
        self.chunks = []        # (dtype, begin, end)
        self._chunk_of = {}     # id(p) -> chunk index
        for dt, n in sizes.items():
            ps = [p for p in self.params if p.dtype == dt]
            k = max(1, min(n_chunks, len(ps)))
            target = n / k
            begin, ci = 0, len(self.chunks)
            for i, p in enumerate(ps):
                _, off, num = self.slot[id(p)]
                end = off + (num + self.ALIGN - 1) // self.ALIGN * self.ALIGN
                self._chunk_of[id(p)] = len(self.chunks)
                last = i == len(ps) - 1
                if last or (end >= target * (len(self.chunks) - ci + 1) and len(self.chunks) - ci < k - 1):
                    self.chunks.append((dt, begin, end))
                    begin = end
        self._members = [0] * len(self.chunks)
        for p in self.params:
            self._members[self._chunk_of[id(p)]] += 1
        self._left = list(self._members)
        self._launched = [False] * len(self.chunks)
        self._written = set()
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._sink_installed = False
        if install_sink:
            self.install()

    # ---- the sink the autograd Functions ask for a gradient buffer
    def install(self):
        from . import functional as VF
        VF.GRAD_SINK = self._sink
        self._sink_installed = True

    def _view(self, p):
        dt, off, n = self.slot[id(p)]
        return self.flat[dt][off:off + n].view(p.shape)

    def _sink(self, key, shape, dtype):
        p = self._by_ptr.get(key)
        if p is None or tuple(p.shape) != tuple(shape) or p.dtype != dtype or id(p) in self._written:
            return None
        self._written.add(id(p))
        return self._view(p)

    # ---- per step
    def zero_grad(self):
        """zero the flat buffers (one memset each) and detach every ``.grad`` so that autograd adopts the slot views"""
        for f in self.flat.values():
            f.zero_()
        for p in self.params:
            p.grad = None
        self._written.clear()
        self._left = list(self._members)
        self._launched = [False] * len(self.chunks)

    def _on_grad(self, p):
        dt, off, n = self.slot[id(p)]
        flat = self.flat[dt]
        g = p.grad
        if g is not None and not (g.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
                                  and g.storage_offset() == off and g.is_contiguous()):
            v = self._view(p)
            v.copy_(g)
            p.grad = v
        ci = self._chunk_of[id(p)]
        self._left[ci] -= 1
        if self._left[ci] == 0 and not self._launched[ci]:
            self._launch(ci)

    def _launch(self, ci):
        self._launched[ci] = True
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        dt, a, b = self.chunks[ci]
        buf = self.flat[dt][a:b]
        if dist.get_backend() == "nccl":
            self._works.append((dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True), None))
        else:   # gloo has no AVG
            self._works.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True), buf))

    def finish(self) -> int:
        """after ``backward()``: send the ranges still waiting for a gradient (their missing slots hold zeros), wait for all
        collectives; returns the number of collectives of this step"""
        for ci in range(len(self.chunks)):
            if not self._launched[ci]:
                self._launch(ci)
        n = len(self._works)
        distributed = dist.is_initialized() and dist.get_world_size() > 1
        present = None
        if distributed and self.unused == "none":
            # which parameters received a gradient on ANY rank (a tiny extra collective, like DDP's unused-parameter bitmap):
            # those get their averaged slot as .grad on every rank, so that all replicas take the same optimizer step
            dev = next(iter(self.flat.values())).device
            present = torch.tensor([0 if p.grad is None else 1 for p in self.params], dtype=torch.int32).to(dev)
            pw = dist.all_reduce(present, op=dist.ReduceOp.SUM, async_op=True)
        for w, buf in self._works:
            w.wait()
            if buf is not None:
                buf.div_(dist.get_world_size())
        self._works.clear()
        if distributed and self.unused == "none":
            pw.wait()
            for p, c in zip(self.params, present.tolist()):
                if c and p.grad is None:
                    p.grad = self._view(p)
        elif distributed:
            for p in self.params:
                if p.grad is None:
                    p.grad = self._view(p)
        return n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
        if self._sink_installed:
            from . import functional as VF
            VF.GRAD_SINK = None
            self._sink_installed = False

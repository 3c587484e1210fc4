"""View-sharded data parallelism: the one multi-GPU strategy the path needs.

Each training view is an independent forward+backward over a read-only replica of the Gaussian
parameters; the only coupling is the SUM of per-Gaussian gradients.  One process per GPU
(torch.distributed, NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests), views assigned
round-robin ``v % world == rank``; every rank accumulates its views' gradients in place into ONE flat
f32 buffer (the ``.grad`` of each parameter is a view into it) and a single all-reduce per step sums
it.  The reference itself is single-GPU (utils/general_utils.py:144) and renders one view per optimizer
step (train.py:121-132); batching views is the data-parallel generalisation.
"""
from __future__ import annotations

from typing import Callable, Iterable, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> list[int]:
    """Indices of the views rank `rank` renders: round-robin, v % world == rank."""
    return list(range(rank, n_views, world))


class FlatGradBuffer:
    """One contiguous f32 gradient buffer; ``p.grad`` of every parameter aliases a slice of it, so
    autograd's in-place accumulation fills the buffer directly and the all-reduce needs no packing."""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params = list(params)
        assert self.params, "no parameters"
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        self.slices = []
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view(p.shape)
            self.slices.append((off, n))
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, async_op: bool = False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    def grads(self):
        return [self.flat[o:o + n].view(p.shape) for (o, n), p in zip(self.slices, self.params)]


class GradArena:
    """Flat f32 buffer whose slices RECEIVE the rasterizer's parameter gradients directly (the C ABI takes
    caller-owned output pointers), so a view needs neither a zero fill, nor autograd's accumulate-add, nor a
    packing copy before the all-reduce.

    `named_params` maps the binding's gradient names ("means3D", "sh", "semantics", "opacities", "scales",
    "rotations", ...) to the parameter tensors.  Use as a context manager around the backward call of EACH view;
    after it, `p.grad` of every parameter is the arena slice.  Entering the context drops any `.grad` that still
    aliases the arena from the previous view (otherwise autograd's AccumulateGrad would add the slice to itself),
    so callers do not have to clear gradients by hand between views.  The arena is installed for the device of its
    own buffer only (diff_gaussian_rasterization._C keeps one per device); every slot starts on a 16-byte boundary
    (the padding floats stay zero)."""

    ALIGN = 4                            # floats: slots start 16-byte aligned whatever P is

    def __init__(self, named_params: dict):
        self.named = dict(named_params)
        ps = list(self.named.values())
        dev = ps[0].device
        pad = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(sum(pad(p.numel()) for p in ps), dtype=torch.float32, device=dev)
        self.slots, off = {}, 0
        for name, p in self.named.items():
            self.slots[name] = self.flat[off:off + p.numel()]
            off += pad(p.numel())
        self.accumulate = False

    def accumulating(self, on: bool = True):
        """`with arena.accumulating(v > 0):` -- the backward inside ADDS its parameter gradients to the arena
        (goi_bwd_out.accumulate) instead of overwriting it: view 0 of a batch overwrites, views 1.. add, and
        one all-reduce follows the last view (BASELINE config 4: 8 views per GPU per step)."""
        self.accumulate = bool(on)
        return self

    def __enter__(self):
        from diff_gaussian_rasterization import _C
        self.clear_grads(only_aliased=True)
        _C.set_grad_arena(self.slots, accumulate=self.accumulate, device=self.flat.device)
        return self

    def __exit__(self, *exc):
        from diff_gaussian_rasterization import _C
        _C.set_grad_arena(None, device=self.flat.device)
        self.accumulate = False
        return False

    def clear_grads(self, only_aliased: bool = False):
        """Set `.grad = None` on the arena's parameters.  Entering the context does this for the gradients that alias
        the arena (i.e. before each view); call it yourself once per step if other code left ordinary `.grad`s."""
        for name, p in self.named.items():
            if p.grad is not None and (not only_aliased or p.grad.data_ptr() == self.slots[name].data_ptr()):
                p.grad = None

    def all_reduce(self, async_op: bool = False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return None


def accumulate_views(views: Iterable, params: Sequence[torch.Tensor], loss_fn: Callable, flat: FlatGradBuffer | None = None,
                     rank: int | None = None, world: int | None = None):
    """Render this rank's share of `views`, back-propagate `loss_fn(view)` for each, and all-reduce.

    loss_fn(view) must return a scalar tensor that depends on `params` (e.g. a pseudo-loss on
    gaussian_renderer.render(view, ...)).  Returns (flat buffer, list of local losses)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    views = list(views)
    flat = flat or FlatGradBuffer(params)
    flat.zero()
    losses = []
    for v in shard_views(len(views), rank, world):
        loss = loss_fn(views[v])
        loss.backward()
        losses.append(loss.detach())
    flat.all_reduce()
    return flat, losses

"""Data-parallel training plumbing (BASELINE configs[4]).

One process per GPU (torchrun); the model is 13.8 MB so it is replicated.  Rank 0's parameters and
BatchNorm buffers are broadcast once; every rank trains on its own shard of the case list with its
own BatchNorm batch statistics (the reference has no SyncBN); after ``loss.backward()`` ONE
all-reduce(SUM) of the flat fp32 gradient (3,453,012 elements) runs over NCCL/NVLink (gloo in the CPU
tests) and the 1/world_size average is folded into the fused Adam step.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_list(items, rank: int, world: int, pad: bool = True):
    """DistributedSampler-style shard: items[rank::world]; with ``pad`` the list is first extended by
    wrapping around so every rank gets the same number of items (equal step counts, no hang)."""
    items = list(items)
    if world <= 1 or not items:
        return items
    if pad and len(items) % world:
        items = items + items[: world - len(items) % world]
    return items[rank::world]


@torch.no_grad()
def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None):
    """Parameters and buffers of rank ``src`` to every rank (one flat broadcast per dtype)."""
    if not is_distributed():
        return
    by_dtype = {}
    for t in list(module.parameters()) + list(module.buffers()):
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            t.copy_(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()


class GradReducer:
    """All-reduces gradients once per step.  With a ``FusedAdam`` the flat gradient buffer is reduced in place
    (SUM; the optimizer applies 1/world); otherwise the gradients are flattened, averaged and copied back."""

    def __init__(self, module: torch.nn.Module, optimizer=None, group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = group
        self.flat = optimizer.flat_grads() if optimizer is not None and hasattr(optimizer, "flat_grads") else []
        self.world = dist.get_world_size(group) if is_distributed() else 1
        if self.flat and optimizer is not None:
            optimizer.grad_scale = 1.0 / self.world

    @torch.no_grad()
    def reduce(self):
        if self.world <= 1:
            return
        if self.flat:
            for g in self.flat:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat /= self.world
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view(g.shape))
            off += g.numel()

"""Multi-GPU sliding-window inference: windows sharded by index, one reduce of the accumulator.

Windows are independent in eval mode, so rank r runs the contiguous block
``shard_range(n_windows, r, world)`` of the MONAI window list into its own zero-initialised fp32
accumulator, and ONE ``reduce(SUM)`` (NCCL over NVLink on GPUs, gloo in the CPU tests) brings the
un-normalised accumulators to rank 0, which divides by the input-independent weight-sum map
(``sliding_window.finalize``).  No other collective touches the data path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import sliding_window as sw


def sharded_sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                                     sigma_scale=0.125, padding_mode="constant", cval=0.0, group=None, dst=0,
                                     label=None, return_mask=False):
    """Same result as ``sliding_window_inference`` on rank ``dst`` (None on the other ranks).

    Every rank passes the same ``inputs`` (the whole volume).  With ``label`` / ``return_mask`` the
    result is ``(probabilities, argmax mask, hard-Dice sums)`` as ``sliding_window.finalize`` returns.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        acc, cnt, lows, img = sw.sliding_window_accumulate(inputs, roi_size, predictor, overlap, mode, sigma_scale,
                                                           padding_mode, cval, sw_batch_size)
        return sw.finalize(acc, cnt, lows, img, label=label, return_mask=return_mask)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    acc, cnt, lows, img = sw.sliding_window_accumulate(inputs, roi_size, predictor, overlap, mode, sigma_scale,
                                                       padding_mode, cval, sw_batch_size, window_shard=(rank, world))
    dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    return sw.finalize(acc, cnt, lows, img, label=label, return_mask=return_mask)


def shard_slab(image_size, roi_size, overlap, rank, world):
    """[x0, x1) along the first spatial axis that covers every window of ``rank``'s shard (one volume,
    batch 1).  Windows are ordered with the first axis slowest, so a contiguous block of window indices
    is a contiguous x slab: a rank only has to receive that part of the input volume (the host->device
    copy of a shard is 1/3 of the volume at 8 ranks for the 384x384x160 / 128^3 benchmark geometry).
    Returns None when the shard owns no window."""
    starts = sw.window_starts(tuple(image_size), tuple(roi_size), overlap)
    lo, hi = sw.shard_range(len(starts), rank, world)
    if hi <= lo:
        return None
    xs = [s[0] for s in starts[lo:hi]]
    return min(xs), min(max(xs) + int(roi_size[0]), int(image_size[0]))

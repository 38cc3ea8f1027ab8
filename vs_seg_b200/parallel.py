"""Multi-GPU sliding-window inference: windows sharded by index over the ranks of one node.

Windows are independent in eval mode, so rank r runs the contiguous block
``shard_range(n_windows, r, world)`` of the MONAI window list.  Two ways to assemble the volume on rank ``dst``:

  * peer blend (default on GPUs, vs_seg_b200.peer): the last kernel of the network blends every window with atomic adds
    straight into ``dst``'s accumulator over NVLink peer memory - the exchange is fused into the compute, there is no
    reduce pass, no per-rank accumulator and no per-rank zero fill; a two-deep arrive/release hand-shake lets the ranks
    run one volume ahead of ``dst``'s finalise;
  * reduce (gloo / CPU tests, foreign predictors, ``VSSEG_SW_PEER=0``): private zero-initialised accumulators and ONE
    ``reduce(SUM)`` of the un-normalised accumulator to ``dst``.

``dst`` divides by the input-independent weight-sum map (``sliding_window.finalize``).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import sliding_window as sw


_PEER: dict = {}
_PEER_DISABLED = False   # set (on every rank together) when CUDA IPC turned out to be unavailable


def _peer_accumulator(shape, device, group, dst):
    """The cached PeerAccumulator for one accumulator shape (collective: every rank creates it at the same call)."""
    from .peer import PeerAccumulator
    key = (tuple(shape), str(device), id(group), dst)
    acc = _PEER.get(key)
    if acc is None:
        for old in _PEER.values():
            old.close()
        _PEER.clear()
        acc = _PEER[key] = PeerAccumulator(shape, device, group, dst)
    return acc


def _use_peer(inputs, predictor):
    return (not _PEER_DISABLED and os.environ.get("VSSEG_SW_PEER", "1") != "0" and inputs.is_cuda and inputs.dim() == 5
            and sw._native_model(predictor) is not None and dist.get_backend() == "nccl")


def sharded_sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                                     sigma_scale=0.125, padding_mode="constant", cval=0.0, group=None, dst=0,
                                     label=None, return_mask=False):
    """Same result as ``sliding_window_inference`` on rank ``dst`` (None on the other ranks).

    Every rank passes the same ``inputs`` (the whole volume, or at least the x slab of its windows, see
    ``shard_slab``).  With ``label`` / ``return_mask`` the result is ``(probabilities, argmax mask, hard-Dice sums)``
    as ``sliding_window.finalize`` returns.

    On one node with the native CUDA path the ranks blend their windows straight into rank ``dst``'s accumulator over
    NVLink peer memory (vs_seg_b200.peer; atomic adds, so the sum order of overlapping windows is unspecified - the
    result agrees with the single-GPU one to fp32 rounding).  ``VSSEG_SW_PEER=0``, CPU tensors and foreign predictors
    use private accumulators and ONE reduce(SUM) of the un-normalised accumulator to ``dst`` instead.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        acc, cnt, lows, img = sw.sliding_window_accumulate(inputs, roi_size, predictor, overlap, mode, sigma_scale,
                                                           padding_mode, cval, sw_batch_size)
        return sw.finalize(acc, cnt, lows, img, label=label, return_mask=return_mask)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pa = None
    if _use_peer(inputs, predictor):
        global _PEER_DISABLED
        from .peer import PeerUnavailable
        model = sw._native_model(predictor)
        nd = inputs.dim() - 2
        roi = (roi_size,) * nd if isinstance(roi_size, int) else tuple(roi_size)
        padded = tuple(max(int(i), int(r) if r and r > 0 else int(i)) for i, r in zip(inputs.shape[2:], roi))
        try:   # collective: every rank maps the destination's buffers, or every rank falls back to the reduce path
            pa = _peer_accumulator((inputs.shape[0], model.out_channels) + padded, inputs.device, group, dst)
        except PeerUnavailable:
            _PEER_DISABLED = True
    if pa is not None:
        _, cnt, lows, img = sw.sliding_window_accumulate(inputs, roi_size, predictor, overlap, mode, sigma_scale,
                                                         padding_mode, cval, sw_batch_size, window_shard=(rank, world),
                                                         peer_acc=lambda shape: pa.begin())
        pa.arrive()
        res = None
        if rank == dst:
            res = sw.finalize(pa.gather(), cnt, lows, img, label=label, return_mask=return_mask)
            pa.release()
        pa.end()
        return res
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

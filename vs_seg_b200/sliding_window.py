"""Sliding-window inference, MONAI-0.4.0-compatible signature, B200-native fast path.

Drop-in for ``monai.inferers.sliding_window_inference`` as imported by the reference at
/root/reference/params/VSparams.py:32 and called at :568-574.  Window geometry and Gaussian
importance map follow the MONAI 0.4.0 algorithm.

Two paths:
  * native (predictor is / wraps a ``UNet2d5_spvPA`` on a CUDA device): every window is read in
    place from the volume by the first conv kernel (no gather copy), runs the fused eval plan,
    and the last conv kernel blends w*logits straight into the fp32 accumulator; the
    input-independent weight-sum map is computed once per geometry and cached; one finalise
    kernel divides (and can emit the argmax mask / Dice sums).  Windows can be sharded over
    ranks (``window_shard``) with a single reduce of the accumulator (vs_seg_b200.parallel).
  * generic (any callable predictor, any device): the same schedule with torch tensor ops, as
    MONAI does; this is host plumbing around a user-supplied predictor, not a kernel fallback.
"""
from __future__ import annotations

import os

import itertools
import math
from typing import Callable, Sequence

import torch
import torch.nn.functional as F

from . import lib as _lib
from .engine import batch_first_enabled
from .tensors import f32view

_IMAP_CACHE: dict = {}
_CNT_CACHE: dict = {}


def _scan_interval(image_size, roi_size, overlap):
    out = []
    for img, roi in zip(image_size, roi_size):
        if roi == img:
            out.append(int(roi))
        else:
            iv = int(roi * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], overlap: float = 0.25):
    """Start corner of every window; first spatial dim slowest, last fastest (MONAI order)."""
    interval = _scan_interval(image_size, roi_size, overlap)
    per_dim = []
    for img, roi, iv in zip(image_size, roi_size, interval):
        num = int(math.ceil(float(img) / iv))
        first = next((d for d in range(num) if d * iv + roi >= img), None)
        n = first + 1 if first is not None else 1
        per_dim.append([i * iv - max(i * iv + roi - img, 0) for i in range(n)])
    return list(itertools.product(*per_dim))


def _gauss_axis(n: int, sigma_scale: float) -> torch.Tensor:
    """Per-axis weights: the erf-integrated Gaussian of MONAI 0.4.0 centred at n//2, zero beyond 4 sigma."""
    sigma = n * sigma_scale
    tail = int(max(float(sigma) * 4.0, 0.5) + 0.5)
    d = torch.arange(n, dtype=torch.float) - (n // 2)
    t = 0.70710678 / abs(float(sigma))
    w = 0.5 * ((t * (d + 0.5)).erf() - (t * (d - 0.5)).erf())
    w = w.clamp(min=0)
    # MONAI blurs a delta at n//2 with a zero-padded "same" correlation of a (2*tail+1)-tap kernel:
    # position i receives kernel[tail + (n//2 - i)], which is symmetric, so w[i] = g(i - n//2).
    w[(d.abs() > tail)] = 0
    return w


def importance_map(roi_size, mode="constant", sigma_scale=0.125, device="cpu") -> torch.Tensor:
    key = (tuple(roi_size), str(mode), float(sigma_scale), str(device))
    if key in _IMAP_CACHE:
        return _IMAP_CACHE[key]
    mode = str(getattr(mode, "value", mode)).lower()
    if mode == "constant":
        m = torch.ones(tuple(roi_size), dtype=torch.float)
    elif mode == "gaussian":
        m = torch.ones((), dtype=torch.float)
        for n in roi_size:  # separable blur of a delta = successive outer products (fp32, same order)
            m = m.unsqueeze(-1) * _gauss_axis(int(n), sigma_scale)
        m = m / m.max()
        m = torch.clamp(m, min=m[m != 0].min().item())
    else:
        raise ValueError(f"unsupported blend mode {mode!r}")
    m = m.float().contiguous().to(device)
    _IMAP_CACHE[key] = m
    return m


def count_map(image_size, roi_size, starts, imap: torch.Tensor) -> torch.Tensor:
    """Sum of importance maps over all windows (input independent; cached per geometry)."""
    key = (tuple(image_size), tuple(roi_size), tuple(starts), imap.data_ptr(), str(imap.device))
    if key in _CNT_CACHE:
        return _CNT_CACHE[key]
    cnt = torch.zeros(tuple(image_size), dtype=torch.float32, device=imap.device)
    for s in starts:  # same accumulation order as the reference loop => same rounding
        sl = tuple(slice(a, a + r) for a, r in zip(s, roi_size))
        cnt[sl] += imap
    _CNT_CACHE[key] = cnt
    return cnt


def _native_model(predictor):
    from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA
    m = getattr(predictor, "native_model", predictor)
    return m if isinstance(m, UNet2d5_spvPA) else None


def _pad_to_roi(inputs, roi_size, padding_mode, cval):
    nd = inputs.dim() - 2
    pad_size = []
    for k in range(inputs.dim() - 1, 1, -1):
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    if any(pad_size):
        mode = str(getattr(padding_mode, "value", padding_mode))
        inputs = F.pad(inputs, pad=pad_size, mode=mode, value=cval)
    lows = [pad_size[(nd - 1 - sp) * 2] for sp in range(nd)]
    return inputs, lows


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of window indices owned by `rank`."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


class _Program:
    """The device-side schedule of one sliding-window geometry: zero-fill of the accumulator plus every launch of
    every window group, captured ONCE in a CUDA graph (the persistent scheduler of the native path).  Windows are
    read in place through relocatable views (vsseg_f32view.indirect): the graph is re-targeted to a new input
    volume by storing its address in one 8-byte device cell, so a volume costs the host three calls (cell store,
    graph launch, finalise) instead of ~250 kernel launches.  The accumulator is owned by the program and reused
    by every call (it is consumed by ``finalize`` / the reduce right after ``run``).

    Windows run in groups (``VSSEG_SW_GROUP``, default 8): the launches that touch a window's own source /
    destination run per window, every other layer once per group (see UNetEvalPlan.window_levels)."""

    def __init__(self, model, inputs, roi_size, jobs, imap, image_size, peer=False):
        dev = inputs.device
        self.peer = bool(peer)
        self.shape, self.stride = tuple(inputs.shape), tuple(inputs.stride())
        group = max(1, int(os.environ.get("VSSEG_SW_GROUP", "8")))
        levels = int(os.environ.get("VSSEG_SW_WINDOW_LEVELS", "1"))
        # a group plan keeps every activation of its windows resident (~400 B per window voxel): keep it
        # within a quarter of the free device memory (large roi sizes, e.g. the reference's 384x384x64)
        free_b, _ = torch.cuda.mem_get_info(dev)
        per_window = 400 * roi_size[0] * roi_size[1] * roi_size[2]
        group = max(1, min(group, int(0.25 * free_b // per_window)))
        self.acc_shape = (inputs.shape[0], model.out_channels) + tuple(image_size)
        # peer mode (multi-GPU): the accumulator belongs to the destination rank (vs_seg_b200.peer); its address
        # comes through a second cell and the blend is atomic.  Otherwise the program owns the accumulator.
        self.acc = None if self.peer else torch.zeros(self.acc_shape, dtype=torch.float32, device=dev)
        self.cell = torch.zeros(2, dtype=torch.int64, device=dev)   # [input base, accumulator base]
        self.imap = imap
        cell = self.cell.data_ptr()
        groups = [jobs[g0:g0 + group] for g0 in range(0, len(jobs), group)]
        # Two window groups whose destination regions are disjoint run concurrently on two streams (VSSEG_SW_STREAMS=1: off)
        # (two plans with their own activation buffers; the latency-bound coarse levels of one group overlap the
        # bandwidth-bound fine levels of the other).  Phases are joined before the next pair starts, so overlapping
        # windows are still blended in a fixed order (no atomics needed).
        self.streams = 2 if (os.environ.get("VSSEG_SW_STREAMS", "2") == "2" and len(groups) >= 2) else 1
        self.side = torch.cuda.Stream(dev) if self.streams == 2 else None
        if self.streams == 1:
            phases = [[i] for i in range(len(groups))]
        elif self.peer:   # atomic blend: any two groups may run concurrently
            phases = [list(range(i, min(i + 2, len(groups)))) for i in range(0, len(groups), 2)]
        else:
            phases = self._pair_groups(groups, roi_size)
        self.calls = []   # (plan, srcs, dsts)
        self.phases = []  # lists of 1 or 2 indices into self.calls
        for ph in phases:
            idx = []
            for slot, gi in enumerate(ph):
                grp = groups[gi]
                srcs = [self._src_view(inputs, b, s, roi_size, cell) for b, s in grp]
                if self.peer:
                    dsts = [self._acc_view(self.acc_shape, b, s, roi_size, cell + 8) for b, s in grp]
                else:
                    dsts = [f32view(self.acc[b:b + 1], s, roi_size) for b, s in grp]
                if len(grp) == 1:
                    call = (model.eval_plan(roi_size, batch=1, device=dev, slot=slot), srcs[0], dsts[0])
                else:
                    call = (model.eval_plan(roi_size, batch=len(grp), device=dev, window_levels=levels, slot=slot,
                                            atomic_out=self.peer), srcs, dsts)
                idx.append(len(self.calls))
                self.calls.append(call)
            self.phases.append(idx)
        self.launches = sum(len(p.steps) for p, _, _ in self.calls)
        self.graph = None
        if os.environ.get("VSSEG_SW_GRAPH", "1") != "0" and self.calls:
            # one eager pass first: lazy module loading, function attributes and the plan caches of the native
            # library must not happen inside a capture
            self.cell[0].fill_(inputs.data_ptr())
            if self.peer:   # the warm-up pass blends into a scratch volume
                scratch = torch.zeros(self.acc_shape, dtype=torch.float32, device=dev)
                self.cell[1].fill_(scratch.data_ptr())
            self._issue()
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            # thread_local: other threads of the process (NCCL's watchdog polls events) must not abort the capture
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self._issue()

    @staticmethod
    def _src_view(inputs, b, start, roi_size, cell):
        """Relocatable view of one window: byte offset from the volume's base address, which the kernels read
        from the device cell at run time."""
        v = f32view(inputs[b:b + 1], start, roi_size)
        v.ptr = v.ptr - inputs.data_ptr()
        v.indirect = cell
        return v

    @staticmethod
    def _acc_view(shape, b, start, roi_size, cell):
        """Relocatable view of one window of a contiguous [B,C,X,Y,Z] accumulator that lives wherever the cell says."""
        _, Cc, X, Y, Z = shape
        sz, sy, sx, sc = 1, Z, Y * Z, X * Y * Z
        off = 4 * (b * Cc * sc + start[0] * sx + start[1] * sy + start[2] * sz)
        return _lib.F32View(off, Cc * sc, sc, sx, sy, sz, 1, Cc, roi_size[0], roi_size[1], roi_size[2], 0, cell)

    @staticmethod
    def _pair_groups(groups, roi_size):
        """Phases of one or two groups; the two groups of a phase blend into disjoint regions of the accumulator."""
        def box(grp):
            bs = [b for b, _ in grp]
            lo = [min(s[d] for _, s in grp) for d in range(3)]
            hi = [max(s[d] for _, s in grp) + roi_size[d] for d in range(3)]
            return min(bs), max(bs), lo, hi

        def disjoint(p, q):
            if p[1] < q[0] or q[1] < p[0]:
                return True
            return any(p[3][d] <= q[2][d] or q[3][d] <= p[2][d] for d in range(3))

        boxes = [box(g) for g in groups]
        left, phases = list(range(len(groups))), []
        while left:
            i = left.pop(0)
            j = next((k for k in left if disjoint(boxes[i], boxes[k])), None)
            if j is None:
                phases.append([i])
            else:
                left.remove(j)
                phases.append([i, j])
        return phases

    def _issue(self):
        if not self.peer:
            self.acc.zero_()
        wptr = self.imap.data_ptr()
        cur = torch.cuda.current_stream(self.cell.device)
        for ph in self.phases:
            if len(ph) == 2:   # fork: the second group of the phase runs on the side stream
                fork = torch.cuda.Event()
                fork.record(cur)
                self.side.wait_event(fork)
                with torch.cuda.stream(self.side):
                    plan, srcs, dsts = self.calls[ph[1]]
                    plan.run(srcs, dsts, wptr, count=False, atomic=self.peer)
                    join = torch.cuda.Event()
                    join.record(self.side)
            plan, srcs, dsts = self.calls[ph[0]]
            plan.run(srcs, dsts, wptr, count=False, atomic=self.peer)
            if len(ph) == 2:
                cur.wait_event(join)

    def run(self, inputs, acc_ptr=None):
        """Accumulate every window of `inputs` (same layout as the volume the program was built for).
        Peer mode: blend (atomically) into the accumulator at device address `acc_ptr`."""
        if tuple(inputs.shape) != self.shape or tuple(inputs.stride()) != self.stride:
            raise ValueError("sliding-window program called with a different volume layout")
        self.cell[0].fill_(inputs.data_ptr())   # the value travels as a kernel argument: no host buffer to race on
        if self.peer:
            self.cell[1].fill_(acc_ptr)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._issue()
        _lib.count_launch(self.launches)
        return self.acc


_PROGRAMS: dict = {}


def _program(model, inputs, roi_size, jobs, imap, image_size, extra, peer=False):
    """Cached _Program for (model weights, volume layout, geometry, shard)."""
    key = (id(model), model._weights_version(), tuple(inputs.shape), tuple(inputs.stride()), str(inputs.device),
           tuple(roi_size), extra, os.environ.get("VSSEG_SW_GROUP", "8"), os.environ.get("VSSEG_SW_WINDOW_LEVELS", "1"),
           os.environ.get("VSSEG_SW_GRAPH", "1"), os.environ.get("VSSEG_SW_STREAMS", "2"), bool(peer),
           batch_first_enabled())
    prog = _PROGRAMS.get(key)
    if prog is None:
        if len(_PROGRAMS) >= 3:   # each program owns an accumulator volume and a captured graph
            _PROGRAMS.clear()
        prog = _PROGRAMS[key] = _Program(model, inputs, roi_size, jobs, imap, image_size, peer)
    return prog


def sliding_window_accumulate(inputs, roi_size, predictor, overlap=0.25, mode="constant", sigma_scale=0.125,
                              padding_mode="constant", cval=0.0, sw_batch_size=1, window_shard=None, peer_acc=None):
    """Steps 1-6 of the MONAI algorithm: returns (acc [B,C,*img], cnt [*img], lows, image_size_).
    peer_acc: callable(shape) -> device address of a shared accumulator (vs_seg_b200.peer): the windows of this
    rank's shard are blended into it with atomics and `acc` is returned as None."""
    nd = inputs.dim() - 2
    if not 0 <= overlap < 1:
        raise AssertionError("overlap must be >= 0 and < 1.")
    image_size_ = list(inputs.shape[2:])
    batch = inputs.shape[0]
    if isinstance(roi_size, int):
        roi_size = (roi_size,) * nd
    roi_size = tuple(int(r) if r and r > 0 else int(i) for r, i in zip(roi_size, image_size_))
    inputs, lows = _pad_to_roi(inputs, roi_size, padding_mode, cval)
    image_size = tuple(inputs.shape[2:])
    starts = window_starts(image_size, roi_size, overlap)
    imap = importance_map(roi_size, mode, sigma_scale, inputs.device)
    cnt = count_map(image_size, roi_size, starts, imap)
    jobs = [(b, s) for b in range(batch) for s in starts]  # batch index outermost (MONAI order)
    if window_shard is not None:
        lo, hi = shard_range(len(jobs), *window_shard)
        jobs = jobs[lo:hi]
    model = _native_model(predictor)
    if model is not None and inputs.is_cuda and nd == 3:
        if model.training:
            raise RuntimeError("the native sliding-window path runs the eval plan: call model.eval() first")
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        extra = (overlap, str(mode), sigma_scale, window_shard)
        if peer_acc is not None:
            prog = _program(model, inputs, roi_size, jobs, imap, image_size, extra, peer=True)
            prog.run(inputs, peer_acc((inputs.shape[0], model.out_channels) + image_size))
            return None, cnt, lows, image_size_
        prog = _program(model, inputs, roi_size, jobs, imap, image_size, extra)
        return prog.run(inputs), cnt, lows, image_size_
    if peer_acc is not None:
        raise _lib.NativeLibraryError("the peer-memory accumulator needs the native CUDA path")
    acc = None
    for g0 in range(0, len(jobs), sw_batch_size):
        grp = jobs[g0:g0 + sw_batch_size]
        sls = [(slice(b, b + 1), slice(None)) + tuple(slice(a, a + r) for a, r in zip(s, roi_size)) for b, s in grp]
        prob = predictor(torch.cat([inputs[sl] for sl in sls]))
        if acc is None:
            acc = torch.zeros((batch, prob.shape[1]) + image_size, dtype=torch.float32, device=inputs.device)
        for j, sl in enumerate(sls):
            acc[sl] += imap * prob[j]
    if acc is None:  # this shard owns no window: still need the accumulator's shape
        n_cls = getattr(getattr(predictor, "native_model", predictor), "out_channels", None)
        if n_cls is None:
            raise ValueError("empty window shard and the predictor does not expose out_channels")
        acc = torch.zeros((batch, n_cls) + image_size, dtype=torch.float32, device=inputs.device)
    return acc, cnt, lows, image_size_


def _label_arg(label_b):
    """(tensor kept alive, pointer, is_u8) of one batch entry's label for vsseg_sw_finalize: uint8 / bool labels
    are read as bytes (4x less traffic than the reference's float labels), anything else as fp32."""
    if label_b.dtype in (torch.uint8, torch.bool):
        t = label_b.contiguous().view(torch.uint8)
        return t, t.data_ptr(), 1
    t = label_b.contiguous().float()
    return t, t.data_ptr(), 0


def finalize(acc, cnt, lows, image_size_, label=None, return_mask=False):
    """Step 7: acc / cnt and crop of the padding.  On CUDA this is one native kernel that can also
    emit the argmax mask and the hard-Dice sums of VSparams.compute_dice_score."""
    crop = (slice(None), slice(None)) + tuple(slice(lo, lo + n) for lo, n in zip(lows, image_size_))
    if not acc.is_cuda:
        out = (acc / cnt)[crop]
        return (out, None, None) if return_mask or label is not None else out
    lib = _lib.load()
    B, Cc = acc.shape[:2]
    n = cnt.numel()
    out = torch.empty_like(acc)
    mask = torch.empty((B, 1) + tuple(acc.shape[2:]), dtype=torch.uint8, device=acc.device) if return_mask else None
    sums = torch.zeros((B, 3), dtype=torch.float64, device=acc.device) if label is not None else None
    if label is not None and tuple(label.shape[2:]) != tuple(acc.shape[2:]):
        raise ValueError("label must have the (un-padded) image shape equal to the accumulator's")
    s = torch.cuda.current_stream(acc.device).cuda_stream
    for b in range(B):
        keep, lptr, u8 = _label_arg(label[b]) if label is not None else (None, None, 0)
        _lib.check(lib.vsseg_sw_finalize(acc[b].data_ptr(), cnt.data_ptr(), out[b].data_ptr(), Cc, n,
                                         mask[b].data_ptr() if mask is not None else None, lptr, u8,
                                         sums[b].data_ptr() if sums is not None else None, s), "sw_finalize")
        _lib.count_launch()
    out = out[crop]
    if return_mask or label is not None:
        return out, (mask[crop[:1] + (slice(None),) + crop[2:]] if mask is not None else None), sums
    return out


def dice_from_sums(sums, smooth=1e-5):
    """Hard foreground Dice per batch entry from the sums of vsseg_sw_finalize: (2I + eps) / (G + P + eps), i.e.
    1 - DiceLoss(include_background=False, to_onehot_y=True) of the one-hot argmax (reference VSparams.py:393-408)."""
    return (2.0 * sums[:, 0] + smooth) / (sums[:, 1] + sums[:, 2] + smooth)


def hard_dice(probabilities: torch.Tensor, label: torch.Tensor, return_mask=False, smooth=1e-5):
    """VSparams.compute_dice_score on the device in ONE launch: argmax over the class dimension, foreground
    Dice sums against `label` (and optionally the uint8 argmax mask).  probabilities: [B,C,*spatial] fp32 CUDA,
    label: [B,1,*spatial] (float / uint8 / bool).  Returns a [B] fp64 device tensor (no host sync)."""
    if not probabilities.is_cuda:
        raise _lib.NativeLibraryError("hard_dice runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    p = probabilities.contiguous().float()
    B, Cc = p.shape[:2]
    n = p[0, 0].numel()
    if label.shape[0] != B or label[0].numel() != n:
        raise ValueError("label must be [B,1,*spatial] with the spatial shape of the probabilities")
    sums = torch.zeros((B, 3), dtype=torch.float64, device=p.device)
    mask = torch.empty((B, 1) + tuple(p.shape[2:]), dtype=torch.uint8, device=p.device) if return_mask else None
    s = torch.cuda.current_stream(p.device).cuda_stream
    for b in range(B):
        keep, lptr, u8 = _label_arg(label[b])
        _lib.check(lib.vsseg_sw_finalize(p[b].data_ptr(), None, None, Cc, n, mask[b].data_ptr() if mask is not None else None,
                                         lptr, u8, sums[b].data_ptr(), s), "sw_finalize")
        _lib.count_launch()
    d = dice_from_sums(sums, smooth)
    return (d, mask) if return_mask else d


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.25, mode="constant", sigma_scale=0.125,
                             padding_mode="constant", cval: float = 0.0, sw_device=None, device=None,
                             *args, **kwargs) -> torch.Tensor:
    """Same signature and result as MONAI 0.4.0's function (reference call: VSparams.py:568-574)."""
    if args or kwargs:
        inner = predictor
        predictor = lambda x: inner(x, *args, **kwargs)  # noqa: E731
        if hasattr(inner, "native_model"):
            predictor.native_model = inner.native_model
    acc, cnt, lows, image_size_ = sliding_window_accumulate(
        inputs, roi_size, predictor, overlap, mode, sigma_scale, padding_mode, cval, sw_batch_size)
    return finalize(acc, cnt, lows, image_size_)

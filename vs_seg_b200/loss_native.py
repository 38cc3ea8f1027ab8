"""Dice_spvPA on the native loss kernels (include/vsseg_b200.h, csrc/vsseg_loss.cu).

Replaces the ~25 elementwise/reduction launches of the reference composition
(/root/reference/params/losses/dice_spvPA.py:238-297) by: a label pyramid (5 max-pools), one
reduction pass per term, a one-thread finalise and one elementwise backward pass per term.  The
softmax / one-hot / hardness weight are evaluated once in registers (the reference computes
softmax and one_hot twice, :282 and :109/:118).  No CPU or torch-op fallback: unsupported
configurations raise.
"""
from __future__ import annotations

import torch

from . import lib as _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _DiceSpvPA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, supervised_attention, hardness_weighting, smooth, *att_maps):
        lib = _lib.load()
        dev = x.device
        if x.dim() != 5 or x.shape[1] != 2:
            raise NotImplementedError("native Dice_spvPA covers the reference configuration: 5-D logits with 2 classes")
        if target.shape[0] != x.shape[0] or target.shape[1] != 1 or tuple(target.shape[2:]) != tuple(x.shape[2:]):
            raise AssertionError(f"ground truth has differing shape ({tuple(target.shape)}) from input ({tuple(x.shape)})")
        s = _stream(dev)
        B = x.shape[0]
        xc, tc = _f32c(x), _f32c(target)
        n = xc[0, 0].numel()
        lam = 0.6 if hardness_weighting else -1.0
        atts = [_f32c(a) for a in att_maps] if supervised_attention else []
        L = len(atts)
        # label pyramid, finest level first (att_maps are coarsest first)
        labels = []
        g = tc
        for level in range(L):
            cur = atts[L - level - 1]
            if tuple(cur.shape) != tuple(g.shape):
                raise AssertionError(f"ground truth has differing shape ({tuple(g.shape)}) from input ({tuple(cur.shape)})")
            labels.append(g)
            if level < L - 1:
                nxt = atts[L - level - 2]
                assert all([a % b == 0 for a, b in zip(cur.shape, nxt.shape)])
                r = [a // b for a, b in zip(cur.shape, nxt.shape)][2:5]
                out = torch.empty((B, 1) + tuple(nxt.shape[2:]), dtype=torch.float32, device=dev)
                _lib.check(lib.vsseg_maxpool3d(g.data_ptr(), out.data_ptr(), B, out.shape[2], out.shape[3], out.shape[4],
                                               r[0], r[1], r[2], s), "maxpool3d")
                g = out
        # rows: [att level 0 (finest) .. L-1] each B rows, then logits B*2 rows
        nrows = L * B + 2 * B
        sums = torch.zeros((nrows, 3), dtype=torch.float64, device=dev)
        # built on the device (a host->device copy of a pageable tensor cannot be captured in a CUDA graph)
        scale = torch.full((nrows,), 1.0 / (2 * B), dtype=torch.float32, device=dev)
        if L:
            scale[:L * B] = 1.0 / (L * B)
        for level in range(L):
            a = atts[L - level - 1]
            _lib.check(lib.vsseg_dice_sums(a.data_ptr(), labels[level].data_ptr(), B, 1, a[0].numel(), -1.0,
                                           sums[level * B].data_ptr(), s), "dice_sums")
        _lib.check(lib.vsseg_dice_sums(xc.data_ptr(), tc.data_ptr(), B, 2, n, lam, sums[L * B].data_ptr(), s), "dice_sums")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        coef = torch.empty((nrows, 2), dtype=torch.float32, device=dev)
        _lib.check(lib.vsseg_dice_finalize(sums.data_ptr(), scale.data_ptr(), nrows, float(smooth), loss.data_ptr(),
                                           coef.data_ptr(), s), "dice_finalize")
        _lib.count_launch(max(L - 1, 0) + L + 2)
        ctx.save_for_backward(xc, tc, coef, *labels)
        ctx.meta = (B, L, n, lam, [tuple(a.shape) for a in att_maps], supervised_attention)
        return loss

    @staticmethod
    def backward(ctx, go):
        lib = _lib.load()
        xc, tc, coef, *labels = ctx.saved_tensors
        B, L, n, lam, att_shapes, supervised = ctx.meta
        dev = xc.device
        s = _stream(dev)
        go = _f32c(go)
        gx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        if gx is not None:
            _lib.check(lib.vsseg_dice_backward(xc.data_ptr(), tc.data_ptr(), B, 2, n, lam, coef[L * B].data_ptr(),
                                               go.data_ptr(), gx.data_ptr(), s), "dice_backward")
            _lib.count_launch()
        gatts = []
        for i, shp in enumerate(att_shapes):  # att_maps order: coarsest first; level = L-1-i
            if not supervised or not ctx.needs_input_grad[5 + i]:
                gatts.append(None)
                continue
            level = L - 1 - i
            ga = torch.empty(shp, dtype=torch.float32, device=dev)
            _lib.check(lib.vsseg_dice_backward(None, labels[level].data_ptr(), B, 1, ga[0].numel(), -1.0,
                                               coef[level * B].data_ptr(), go.data_ptr(), ga.data_ptr(), s),
                       "dice_backward")
            _lib.count_launch()
            gatts.append(ga)
        return (gx, None, None, None, None, *gatts)


def dice_spvpa_native(x, att_maps, target, supervised_attention=True, hardness_weighting=True, smooth=1e-5):
    """Dice_spvPA.forward((x, att_maps), target) on CUDA tensors, differentiable w.r.t. x and att_maps."""
    if not x.is_cuda:
        raise _lib.NativeLibraryError("dice_spvpa_native needs CUDA tensors (no CPU fallback)")
    return _DiceSpvPA.apply(x, target, bool(supervised_attention), bool(hardness_weighting), float(smooth), *att_maps)


class _DiceSums(torch.autograd.Function):
    """sums[b][c] = (sum w t p, sum w t', sum w p') of DiceLoss (reference dice_spvPA.py:133-149) in one native pass,
    differentiable w.r.t. the prediction (one native elementwise pass backward)."""

    @staticmethod
    def forward(ctx, pred, target, weight, act, onehot, squared):
        lib = _lib.load()
        B, C = pred.shape[:2]
        n = pred[0, 0].numel()
        pc, tc = _f32c(pred), _f32c(target)
        wc = _f32c(weight) if weight is not None else None
        sums = torch.zeros((B, 8, 3), dtype=torch.float64, device=pred.device)
        _lib.check(lib.vsseg_dice_general_sums(pc.data_ptr(), tc.data_ptr(), wc.data_ptr() if wc is not None else None, B, C, n,
                                               act, int(onehot), int(squared), sums.data_ptr(), _stream(pred.device)),
                   "dice_general_sums")
        _lib.count_launch()
        ctx.save_for_backward(pc, tc, wc if wc is not None else pc.new_empty(0))
        ctx.meta = (B, C, n, act, int(onehot), int(squared), wc is not None, tuple(pred.shape))
        return sums[:, :C].float()

    @staticmethod
    def backward(ctx, gs):
        if not ctx.needs_input_grad[0]:
            return (None,) * 6
        lib = _lib.load()
        pc, tc, wc = ctx.saved_tensors
        B, C, n, act, onehot, squared, has_w, shape = ctx.meta
        g8 = torch.zeros((B, 8, 3), dtype=torch.float32, device=pc.device)
        g8[:, :C] = gs
        grad = torch.empty(shape, dtype=torch.float32, device=pc.device)
        _lib.check(lib.vsseg_dice_general_backward(pc.data_ptr(), tc.data_ptr(), wc.data_ptr() if has_w else None, B, C, n, act,
                                                   onehot, squared, g8.data_ptr(), grad.data_ptr(), _stream(pc.device)),
                   "dice_general_backward")
        _lib.count_launch()
        return grad, None, None, None, None, None


def dice_loss_native(mod, input, target, smooth=1e-5):
    """DiceLoss.forward (reference dice_spvPA.py:90-167) on CUDA tensors: the voxel work is one native reduction
    (and one native backward pass); include_background / jaccard / reduction act on the B x C sums."""
    import warnings
    if not input.is_cuda:
        raise _lib.NativeLibraryError("dice_loss_native needs CUDA tensors (no CPU fallback)")
    n_pred_ch = input.shape[1]
    if n_pred_ch > 8:
        raise NotImplementedError("native DiceLoss covers up to 8 channels")
    act = 0
    if mod.sigmoid:
        act = 1
    if mod.softmax:
        if n_pred_ch == 1:
            warnings.warn("single channel prediction, `softmax=True` ignored.")
        else:
            act = 2
    if mod.other_act is not None:
        input = mod.other_act(input)   # the caller's own callable; the reduction below stays native
    onehot = False
    if mod.to_onehot_y:
        if n_pred_ch == 1:
            warnings.warn("single channel prediction, `to_onehot_y=True` ignored.")
        else:
            onehot = True
    if not mod.include_background and n_pred_ch == 1:
        warnings.warn("single channel prediction, `include_background=False` ignored.")
    want = (input.shape[0], 1) + tuple(input.shape[2:]) if onehot else tuple(input.shape)
    if tuple(target.shape) != want:
        raise AssertionError(f"ground truth has differing shape ({tuple(target.shape)}) from input ({tuple(input.shape)})")
    w = mod.hardness_weight
    if w is not None:
        if w.requires_grad:
            raise NotImplementedError("native DiceLoss treats hardness_weight as a constant; the differentiable hardness "
                                      "weight is the fused Dice_spvPA path")
        w = w.expand_as(input)
    sums = _DiceSums.apply(input, target, w, act, onehot, bool(mod.squared_pred))
    if not mod.include_background and n_pred_ch > 1:
        sums = sums[:, 1:]
    inter, ground_o, pred_o = sums.unbind(-1)
    denom = ground_o + pred_o
    if mod.jaccard:
        denom = 2.0 * (denom - inter)
    f = 1.0 - (2.0 * inter + smooth) / (denom + smooth)
    if mod.reduction == "mean":
        return torch.mean(f)
    if mod.reduction == "sum":
        return torch.sum(f)
    if mod.reduction == "none":
        return f
    raise ValueError(f'Unsupported reduction: {mod.reduction}, available options are ["mean", "sum", "none"].')

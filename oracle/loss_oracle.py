"""Oracle: Dice_spvPA / DiceLoss / Dice metric in plain torch fp32 or fp64 (test infrastructure only).

Follows /root/reference/params/losses/dice_spvPA.py:90-167 (DiceLoss.forward) and :238-297
(Dice_spvPA.forward), and params/VSparams.py:393-408 (compute_dice_score).  Gradients come
from torch autograd, so ``loss.backward()`` on these functions is the gradient oracle.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def one_hot(labels, num_classes):
    """MONAI ``one_hot`` for [B,1,...] float labels (used at dice_spvPA.py:118,282)."""
    shape = list(labels.shape)
    shape[1] = num_classes
    out = torch.zeros(shape, dtype=labels.dtype, device=labels.device)
    return out.scatter_(1, labels.long(), 1)


def dice_loss(pred, target, weight=None, smooth=1e-5):
    """DiceLoss core with every flag False (dice_spvPA.py:133-159), reduction 'mean'."""
    assert pred.shape == target.shape
    axes = list(range(2, pred.dim()))
    if weight is not None:
        inter = torch.sum(weight * target * pred, dim=axes)
        g = torch.sum(weight * target, dim=axes)
        p = torch.sum(weight * pred, dim=axes)
    else:
        inter = torch.sum(target * pred, dim=axes)
        g = torch.sum(target, dim=axes)
        p = torch.sum(pred, dim=axes)
    f = 1.0 - (2.0 * inter + smooth) / (g + p + smooth)
    return f.mean()


def dice_spvpa_loss(x, att_maps, target, supervised_attention=True, hardness_weighting=True,
                    smooth=1e-5, hardness_lambda=0.6):
    """Dice_spvPA.forward((x, att_maps), target) (dice_spvPA.py:238-297)."""
    total_att = x.new_zeros(())
    if supervised_attention:
        n = len(att_maps)
        g = target
        for level in range(n):
            a = att_maps[n - level - 1]
            total_att = total_att + dice_loss(a, g, smooth=smooth) / n
            if level < n - 1:
                nxt = att_maps[n - level - 2]
                assert all(c % m == 0 for c, m in zip(a.shape, nxt.shape))
                ratio = [c // m for c, m in zip(a.shape, nxt.shape)][2:5]
                g = F.max_pool3d(g, kernel_size=ratio, stride=ratio)
    w = None
    t1h = one_hot(target, x.shape[1])
    p = torch.softmax(x, dim=1)
    if hardness_weighting:
        w = hardness_lambda * torch.abs(p - t1h) + (1.0 - hardness_lambda)  # NOT detached (:281-283)
    return total_att + dice_loss(p, t1h, weight=w, smooth=smooth)


def dice_score(probabilities, label, smooth=1e-5):
    """VSparams.compute_dice_score (VSparams.py:393-408): hard foreground Dice."""
    n = probabilities.shape[1]
    y = one_hot(torch.argmax(probabilities, dim=1, keepdim=True).to(label.dtype), n)
    t = one_hot(label, n)
    return 1.0 - dice_loss(y[:, 1:], t[:, 1:], smooth=smooth)

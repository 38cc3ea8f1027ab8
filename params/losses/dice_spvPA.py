"""Dice losses with the reference's names, signatures and error behaviour
(/root/reference/params/losses/dice_spvPA.py:23-297; aliases ``dice = Dice = DiceLoss`` :639).

``Dice_spvPA`` = mean of the per-level attention-map Dice losses against the max-pooled label plus
the hardness-weighted 2-class soft Dice.  On CUDA tensors the reference configuration
(softmax + one-hot, 2 classes, mean reduction) runs as ONE fused forward reduction kernel and ONE
fused backward kernel (vs_seg_b200.loss_native) instead of ~25 elementwise/reduction launches; CPU
tensors use the torch composition below (host plumbing, BASELINE config 1).
"""
import warnings
from typing import Callable, Optional, Union

import torch
import torch.nn.functional as F
from torch.nn.modules.loss import _Loss

from vs_seg_b200.compat import LossReduction, one_hot


class DiceLoss(_Loss):
    """Soft Dice between ``input`` (BNH[WD]) and ``target`` (B1H[WD] or BNH[WD]), optionally voxel-weighted
    by ``hardness_weight``."""

    def __init__(
        self,
        include_background: bool = True,
        to_onehot_y: bool = False,
        sigmoid: bool = False,
        softmax: bool = False,
        other_act: Optional[Callable] = None,
        squared_pred: bool = False,
        jaccard: bool = False,
        hardness_weight=None,
        reduction: Union[LossReduction, str] = LossReduction.MEAN,
    ) -> None:
        super().__init__(reduction=LossReduction(reduction).value)
        if other_act is not None and not callable(other_act):
            raise TypeError(f"other_act must be None or callable but is {type(other_act).__name__}.")
        if int(sigmoid) + int(softmax) + int(other_act is not None) > 1:
            raise ValueError("Incompatible values: more than 1 of [sigmoid=True, softmax=True, other_act is not None].")
        self.include_background = include_background
        self.to_onehot_y = to_onehot_y
        self.sigmoid = sigmoid
        self.softmax = softmax
        self.other_act = other_act
        self.squared_pred = squared_pred
        self.jaccard = jaccard
        self.hardness_weight = hardness_weight

    def forward(self, input: torch.Tensor, target: torch.Tensor, smooth: float = 1e-5) -> torch.Tensor:
        if input.is_cuda:   # one native reduction pass (+ one native backward pass), every constructor flag
            from vs_seg_b200.loss_native import dice_loss_native
            return dice_loss_native(self, input, target, smooth)
        n_pred_ch = input.shape[1]
        if self.sigmoid:
            input = torch.sigmoid(input)
        if self.softmax:
            if n_pred_ch == 1:
                warnings.warn("single channel prediction, `softmax=True` ignored.")
            else:
                input = torch.softmax(input, dim=1)
        if self.other_act is not None:
            input = self.other_act(input)
        if self.to_onehot_y:
            if n_pred_ch == 1:
                warnings.warn("single channel prediction, `to_onehot_y=True` ignored.")
            else:
                target = one_hot(target, num_classes=n_pred_ch)
        if not self.include_background:
            if n_pred_ch == 1:
                warnings.warn("single channel prediction, `include_background=False` ignored.")
            else:
                target, input = target[:, 1:], input[:, 1:]
        assert (
            target.shape == input.shape
        ), f"ground truth has differing shape ({target.shape}) from input ({input.shape})"

        axes = list(range(2, len(input.shape)))  # spatial dims only
        w = self.hardness_weight
        inter = torch.sum(target * input if w is None else w * target * input, dim=axes)
        if self.squared_pred:
            target, input = torch.pow(target, 2), torch.pow(input, 2)
        ground_o = torch.sum(target if w is None else w * target, dim=axes)
        pred_o = torch.sum(input if w is None else w * input, dim=axes)
        denom = ground_o + pred_o
        if self.jaccard:
            denom = 2.0 * (denom - inter)
        f = 1.0 - (2.0 * inter + smooth) / (denom + smooth)
        if self.reduction == LossReduction.MEAN.value:
            return torch.mean(f)
        if self.reduction == LossReduction.SUM.value:
            return torch.sum(f)
        if self.reduction == LossReduction.NONE.value:
            return f
        raise ValueError(f'Unsupported reduction: {self.reduction}, available options are ["mean", "sum", "none"].')


class Dice_spvPA(_Loss):
    """loss = (1/L) sum_l Dice(att_l, maxpool_l(target)) + Dice_w(softmax(x), onehot(target)),
    w = 0.6*|softmax(x) - onehot(target)| + 0.4 (gradient flows through w)."""

    def __init__(
        self,
        include_background: bool = True,
        to_onehot_y: bool = False,
        sigmoid: bool = False,
        softmax: bool = False,
        other_act: Optional[Callable] = None,
        squared_pred: bool = False,
        jaccard: bool = False,
        reduction: Union[LossReduction, str] = LossReduction.MEAN,
        supervised_attention=True,
        hardness_weighting=True,
    ) -> None:
        super().__init__(reduction=LossReduction(reduction).value)
        if other_act is not None and not callable(other_act):
            raise TypeError(f"other_act must be None or callable but is {type(other_act).__name__}.")
        if int(sigmoid) + int(softmax) + int(other_act is not None) > 1:
            raise ValueError("Incompatible values: more than 1 of [sigmoid=True, softmax=True, other_act is not None].")
        self.include_background = include_background
        self.to_onehot_y = to_onehot_y
        self.sigmoid = sigmoid
        self.softmax = softmax
        self.other_act = other_act
        self.squared_pred = squared_pred
        self.jaccard = jaccard
        self.supervised_attention = supervised_attention
        self.hardness_weighting = hardness_weighting

    def forward(self, input, target: torch.Tensor, smooth: float = 1e-5) -> torch.Tensor:
        x, att_maps = input
        if x.is_cuda:
            # like the reference's forward (dice_spvPA.py:250-297), which builds its two DiceLoss instances with fixed
            # flags and ignores the constructor's include_background / to_onehot_y / ... arguments
            from vs_seg_b200.loss_native import dice_spvpa_native
            return dice_spvpa_native(x, att_maps, target, self.supervised_attention, self.hardness_weighting, smooth)

        single = Dice(to_onehot_y=False, softmax=False)
        total_att_loss = 0
        if self.supervised_attention:
            n_levels = len(att_maps)
            g_l = target
            for level in range(n_levels):
                cur = att_maps[n_levels - level - 1]  # finest first
                total_att_loss = total_att_loss + 1 / n_levels * single(cur, g_l)
                if level < n_levels - 1:
                    nxt = att_maps[n_levels - level - 2]
                    assert all([a % b == 0 for a, b in zip(cur.shape, nxt.shape)])
                    ratio = [a // b for a, b in zip(cur.shape, nxt.shape)][2:5]
                    g_l = F.max_pool3d(g_l, kernel_size=ratio, stride=ratio)
        hardness_weight = None
        if self.hardness_weighting:
            lam = 0.6
            hardness_weight = lam * abs(torch.softmax(x, dim=1) - one_hot(target, num_classes=x.shape[1])) + (1.0 - lam)
        multi = Dice(to_onehot_y=True, softmax=True, hardness_weight=hardness_weight)
        return total_att_loss + multi(x, target)


dice = Dice = DiceLoss

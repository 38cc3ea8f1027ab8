"""Convolution / ResidualUnit blocks with the reference's constructor signatures and state_dict keys
(/root/reference/params/networks/blocks/convolutions.py:62-78, :189-204), executed on B200 by the
fused native kernels.

The torch submodules (conv, norm, dropout, act) are parameter containers, so ``state_dict()`` is
identical to the reference's.  On a CUDA tensor ``forward`` never runs them: one fused
conv+BN+act(+residual) kernel per block runs through the C ABI (include/vsseg_b200.h).  On a CPU
tensor the containers execute as plain torch modules — host plumbing for BASELINE config 1
(`VS_train.py --debug` without a GPU), not a fallback of the CUDA path: a CUDA tensor without the
native library raises.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from vs_seg_b200.compat import Act, Conv, Dropout, Norm, same_padding, split_args


def _triple(v, n):
    if isinstance(v, (int, np.integer)):
        return (int(v),) * n
    return tuple(int(a) for a in v)


class Convolution(nn.Sequential):
    """(Conv|ConvTrans) -> Norm -> (Dropout) -> (Act); ``conv_only`` keeps the convolution alone."""

    def __init__(
        self,
        dimensions: int,
        in_channels: int,
        out_channels: int,
        strides: Union[Sequence[int], int] = 1,
        kernel_size: Union[Sequence[int], int] = 3,
        act: Optional[Union[Tuple, str]] = Act.PRELU,
        norm: Optional[Union[Tuple, str]] = Norm.INSTANCE,
        dropout: Optional[Union[Tuple, str, float]] = None,
        dropout_dim: int = 1,
        dilation: Union[Sequence[int], int] = 1,
        groups: int = 1,
        bias: bool = True,
        conv_only: bool = False,
        is_transposed: bool = False,
    ) -> None:
        super().__init__()
        self.dimensions = dimensions
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.is_transposed = is_transposed
        padding = same_padding(kernel_size, dilation)
        conv_cls = Conv[Conv.CONVTRANS if is_transposed else Conv.CONV, dimensions]
        kwargs = dict(kernel_size=kernel_size, stride=strides, padding=padding, groups=groups, bias=bias,
                      dilation=dilation)
        if is_transposed:
            # output = input * stride in every dimension (reference :114-123)
            opad = np.array(strides) + 2 * np.array(padding) - np.array(dilation) * (np.array(kernel_size) - 1) - 1
            kwargs["output_padding"] = int(opad) if opad.size == 1 else tuple(int(v) for v in opad)
        self.add_module("conv", conv_cls(in_channels, out_channels, **kwargs))
        self._act_name = None
        self._norm_name = None
        if not conv_only:
            if norm is not None:
                name, nargs = split_args(norm)
                self._norm_name = str(name).upper()
                self.add_module("norm", Norm[name, dimensions](out_channels, **nargs))
            if dropout:
                if isinstance(dropout, (int, float)):
                    dname, dargs = Dropout.DROPOUT, {"p": dropout}
                else:
                    dname, dargs = split_args(dropout)
                if dropout_dim > dimensions:
                    raise ValueError(
                        f"dropout_dim should be no larger than dimensions, got dropout_dim={dropout_dim} "
                        f"and dimensions={dimensions}.")
                self.add_module("dropout", Dropout[dname, dropout_dim](**dargs))
            if act is not None:
                name, aargs = split_args(act)
                self._act_name = str(name).upper()
                self.add_module("act", Act[name](**aargs))

    # ---- native path -------------------------------------------------------------------
    def _native_supported(self):
        c = self.conv
        return (self.dimensions == 3 and c.groups == 1 and tuple(c.dilation) == (1, 1, 1)
                and all(k in (1, 3) for k in c.kernel_size) and all(s in (1, 2) for s in c.stride)
                and self._norm_name in (None, "BATCH") and self._act_name in (None, "PRELU", "RELU", "SIGMOID")
                and (self._act_name != "PRELU" or self.act.weight.numel() == 1)
                and (not self.is_transposed or all(s == 1 or k == 3 for s, k in zip(c.stride, c.kernel_size))))

    def _native_forward(self, x, residual: Optional[torch.Tensor] = None):
        """Eval-mode fused block on CUDA: pack -> conv+BN+act(+residual) -> unpack."""
        from vs_seg_b200.engine import conv_block_ncdhw

        if self.training and residual is None and self._norm_name == "BATCH" and self._act_name == "PRELU":
            # train mode (batch-statistics BatchNorm, dropout, autograd) on the native training kernels
            from vs_seg_b200.training import block_train_forward
            return block_train_forward(self, x)
        if self.training and (self._norm_name is not None or "dropout" in self._modules):
            raise NotImplementedError(
                "this train-mode Convolution configuration has no native kernel (covered: BatchNorm + PReLU blocks, "
                "channels % 8 == 0); call .eval() for the fused inference block (there is no eager CUDA fallback)")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # the fused inference kernel runs outside autograd: its output would carry no grad_fn and the
            # parameters would silently receive no gradient
            raise NotImplementedError(
                "a standalone Convolution block in train mode with gradients enabled has no native backward; "
                "use torch.no_grad() / .eval(), or train through UNet2d5_spvPA")
        if not self._native_supported():
            raise NotImplementedError("this Convolution configuration has no native sm_100a kernel")
        c = self.conv
        sd = {"conv." + k: v.detach() for k, v in c.state_dict().items()}
        if self._norm_name:
            sd.update({"norm." + k: v.detach() for k, v in self.norm.state_dict().items()})
        if self._act_name == "PRELU":
            sd["act.weight"] = self.act.weight.detach()
        act = {None: "none", "PRELU": "prelu", "RELU": "relu", "SIGMOID": "sigmoid"}[self._act_name]
        return conv_block_ncdhw(x, sd, c.kernel_size, c.stride, self.is_transposed, self._norm_name is not None, act,
                                residual)

    def forward(self, x):
        if x.is_cuda:
            return self._native_forward(x)
        return super().forward(x)


class ResidualUnit(nn.Module):
    """``subunits`` Convolution blocks plus a shortcut conv, summed with no activation after the sum
    (reference convolutions.py:209-255)."""

    def __init__(
        self,
        dimensions: int,
        in_channels: int,
        out_channels: int,
        strides: Union[Sequence[int], int] = 1,
        kernel_size: Union[Sequence[int], int] = 3,
        subunits: int = 2,
        act: Optional[Union[Tuple, str]] = Act.PRELU,
        norm: Optional[Union[Tuple, str]] = Norm.INSTANCE,
        dropout: Optional[Union[Tuple, str, float]] = None,
        dropout_dim: int = 1,
        dilation: Union[Sequence[int], int] = 1,
        bias: bool = True,
        last_conv_only: bool = False,
    ) -> None:
        super().__init__()
        self.dimensions = dimensions
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.conv = nn.Sequential()
        self.residual = nn.Identity()
        padding = same_padding(kernel_size, dilation)
        cin, stride = in_channels, strides
        subunits = max(1, subunits)
        for su in range(subunits):
            unit = Convolution(dimensions, cin, out_channels, strides=stride, kernel_size=kernel_size, act=act,
                               norm=norm, dropout=dropout, dropout_dim=dropout_dim, dilation=dilation, bias=bias,
                               conv_only=last_conv_only and su == subunits - 1)
            self.conv.add_module(f"unit{su:d}", unit)
            cin, stride = out_channels, 1  # later units keep channels and resolution
        if np.prod(strides) != 1 or in_channels != out_channels:
            rk, rp = kernel_size, padding
            if np.prod(strides) == 1:  # channel change only: 1x1x1, no padding
                rk, rp = 1, 0
            self.residual = Conv[Conv.CONV, dimensions](in_channels, out_channels, rk, strides, rp, bias=bias)

    def _native_forward(self, x):
        from vs_seg_b200.engine import native_shortcut
        if self.training:   # batch-statistics BatchNorm / dropout / autograd: the native training tape of the unit
            from vs_seg_b200.training import block_train_forward
            return block_train_forward(self, x)
        res = x if isinstance(self.residual, nn.Identity) else native_shortcut(self.residual, x)
        cx = x
        units = list(self.conv.children())
        for i, u in enumerate(units):
            cx = u._native_forward(cx, residual=res if i == len(units) - 1 else None)
        return cx

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.is_cuda:
            return self._native_forward(x)
        return self.conv(x) + self.residual(x)

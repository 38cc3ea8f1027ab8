"""Spatial attention gate with the reference's class names and signatures
(/root/reference/params/networks/blocks/attentionblock.py:6-47).

AttentionBlock1: conv(C -> C/2, k) + ReLU, conv(C/2 -> 1, k) + Sigmoid; returns (att, x).
AttentionBlock2: x * att (broadcast over channels) + x.
On CUDA both run on the native kernels (conv blocks + one fused gate kernel).
"""
import ctypes as C

import torch

from params.networks.blocks.convolutions import Convolution
from vs_seg_b200.compat import Act


class AttentionBlock1(torch.nn.Module):
    def __init__(self, dimensions, in_channels, out_channels, kernel_size, norm, dropout):
        super(AttentionBlock1, self).__init__()
        self.in_channels = in_channels
        self.conv1 = Convolution(dimensions, in_channels, in_channels // 2, strides=1, kernel_size=kernel_size,
                                 act=Act.RELU, norm=norm, dropout=None)
        self.conv2 = Convolution(dimensions, in_channels // 2, out_channels=1, strides=1, kernel_size=kernel_size,
                                 act=Act.SIGMOID, norm=norm, dropout=None)

    def forward(self, x):
        att = self.conv2(self.conv1(x))
        return att, x


class AttentionBlock2(torch.nn.Module):
    def __init__(self, dimensions, in_channels, out_channels, kernel_size, norm, dropout):
        super(AttentionBlock2, self).__init__()
        self.in_channels = in_channels

    def forward(self, input_tuple):
        att, x = input_tuple
        if not x.is_cuda:
            return att.repeat([1, self.in_channels, 1, 1, 1]) * x + x
        from vs_seg_b200 import lib as _lib
        from vs_seg_b200.tensors import Act8Buffer, f32view
        lib = _lib.load()
        B, c = x.shape[0], x.shape[1]
        c8 = (c + 7) // 8 * 8
        xin = x.float()
        if c8 != c:
            xin = torch.nn.functional.pad(xin, (0, 0, 0, 0, 0, 0, 0, c8 - c))
        buf = Act8Buffer(B, c8, *x.shape[2:], x.device).from_ncdhw(xin)
        v, a = buf.view(), f32view(att.float().contiguous())
        _lib.check(lib.vsseg_att_gate(C.byref(v), C.byref(a), C.byref(v),
                                      torch.cuda.current_stream(x.device).cuda_stream), "att_gate")
        _lib.count_launch()
        return buf.to_ncdhw(0, c8)[:, :c].contiguous()

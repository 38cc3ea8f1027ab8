"""Device tensors in the layouts the native kernels use (torch owns the memory).

act8: channel-blocked split-bf16, see include/vsseg_b200.h.  A buffer is one torch tensor
[2 (hi/lo), B, C/8, X, Y, Z, 8] bf16; channel-range views (torch.cat of a skip connection,
reference unet2d5_spvPA.py:89) are just pointer offsets.
"""
from __future__ import annotations

import torch

from . import lib as _lib


def _stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


class Act8Buffer:
    def __init__(self, B, C, X, Y, Z, device):
        if C % 8:
            raise ValueError(f"act8 needs channels % 8 == 0, got {C}")
        self.B, self.C, self.X, self.Y, self.Z = int(B), int(C), int(X), int(Y), int(Z)
        self.t = torch.empty((2, self.B, self.C // 8, self.X, self.Y, self.Z, 8), dtype=torch.bfloat16, device=device)
        self.device = self.t.device

    @property
    def nvox(self):
        return self.X * self.Y * self.Z

    def view(self, c0=0, C=None, b0=0, nb=None) -> _lib.Act8:
        """Channel range [c0, c0+C) of batch entries [b0, b0+nb) (pointer offsets only)."""
        C = self.C - c0 if C is None else C
        nb = self.B - b0 if nb is None else nb
        if c0 % 8 or C % 8 or c0 + C > self.C:
            raise ValueError("act8 views must be aligned to 8 channels")
        if b0 < 0 or nb < 1 or b0 + nb > self.B:
            raise ValueError("act8 view outside the batch")
        return _lib.Act8(self.t.data_ptr() + ((c0 // 8) * self.nvox * 8 + b0 * self.C * self.nvox) * 2,
                         self.B * self.C * self.nvox, self.C * self.nvox, nb, C, self.X, self.Y, self.Z)

    def to_ncdhw(self, c0=0, C=None) -> torch.Tensor:
        """fp32 [B,C,X,Y,Z] copy through the native unpack kernel."""
        C = self.C - c0 if C is None else C
        out = torch.empty((self.B, C, self.X, self.Y, self.Z), dtype=torch.float32, device=self.device)
        v, o = self.view(c0, C), f32view(out)
        _lib.check(_lib.load().vsseg_unpack_act8(v, o, _stream_ptr(self.device)), "unpack_act8")
        _lib.count_launch()
        return out

    def from_ncdhw(self, x: torch.Tensor, c0=0):
        x = x.contiguous().float()
        v, s = self.view(c0, x.shape[1]), f32view(x)
        _lib.check(_lib.load().vsseg_pack_act8(s, v, _stream_ptr(self.device)), "pack_act8")
        _lib.count_launch()
        return self


def f32view(t: torch.Tensor, offset=(0, 0, 0), size=None, cell: int | None = None) -> _lib.F32View:
    """View of a [B,C,X,Y,Z] fp32 tensor region starting at spatial `offset` with spatial `size`.

    cell: device address of an 8-byte cell that will hold ``t.data_ptr()`` at run time; the view then stores
    only the region's byte offset (relocatable view, include/vsseg_b200.h) so that launches captured in a
    CUDA graph follow whatever volume the cell points to."""
    if t.dtype != torch.float32 or t.dim() != 5:
        raise ValueError("f32view needs a 5-D float32 tensor")
    B, Cc = t.shape[0], t.shape[1]
    sb, sc, sx, sy, sz = t.stride()
    ox, oy, oz = offset
    X, Y, Z = size if size is not None else (t.shape[2] - ox, t.shape[3] - oy, t.shape[4] - oz)
    if ox < 0 or oy < 0 or oz < 0 or ox + X > t.shape[2] or oy + Y > t.shape[3] or oz + Z > t.shape[4]:
        raise ValueError("f32view region outside the tensor")
    off = 4 * (ox * sx + oy * sy + oz * sz)
    if cell is not None:
        return _lib.F32View(off, sb, sc, sx, sy, sz, B, Cc, X, Y, Z, 0, cell)
    return _lib.F32View(t.data_ptr() + off, sb, sc, sx, sy, sz, B, Cc, X, Y, Z, 0, None)

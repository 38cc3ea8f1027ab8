"""Eval-mode execution plan of UNet2d5_spvPA on the native kernels.

The plan is derived from a reference-format ``state_dict`` (the source of truth); folded /
packed weights are caches.  It mirrors the recursion of the reference builder
(/root/reference/params/networks/nets/unet2d5_spvPA.py:56-89) but executes it as a flat list
of fused kernel launches through the C ABI:

  * Conv -> BatchNorm(eval) -> Dropout(eval = identity) -> PReLU is ONE launch
    (convolutions.py:148-156), BN and bias folded into a per-channel scale/shift;
  * the ResidualUnit sum is applied in the last conv's epilogue (convolutions.py:252-255);
  * torch.cat of the skip connection is free: encoder and upsample write disjoint channel
    ranges of one act8 buffer (MONAI SkipConnection, unet2d5_spvPA.py:89);
  * the attention gate x*(1+att) runs in place on that buffer (attentionblock.py:44-47);
  * the top ResidualUnit (conv_only + 1x1x1 shortcut, both linear) is a single conv whose
    centre tap absorbs the shortcut, and can blend straight into the sliding-window
    accumulator (MONAI sliding_window_inference step 6).
"""
from __future__ import annotations

import ctypes as C
import os
import math

import torch

from . import lib as _lib
from .tensors import Act8Buffer, f32view


def _round_up(v, m):
    return (v + m - 1) // m * m


def pack_conv_weight(w: torch.Tensor, transposed: bool, cout_pad: int | None = None) -> torch.Tensor:
    """torch conv weight -> fp32 [taps][Cin][CoutPad] (tap = (tx*ky+ty)*kz+tz)."""
    if transposed:  # ConvTranspose3d: [Cin, Cout, kx, ky, kz]
        p = w.permute(2, 3, 4, 0, 1)
    else:  # Conv3d: [Cout, Cin, kx, ky, kz]
        p = w.permute(2, 3, 4, 1, 0)
    kx, ky, kz, cin, cout = p.shape
    p = p.reshape(kx * ky * kz, cin, cout).float()
    if cout_pad is not None and cout_pad != cout:
        p = torch.nn.functional.pad(p, (0, cout_pad - cout))
    return p.contiguous()


def _split_planes(w: torch.Tensor) -> torch.Tensor:
    w = w.float()
    hi = w.bfloat16()
    lo = (w - hi.float()).bfloat16()
    return torch.stack([hi, lo])


def pack_conv_weight_tc(w: torch.Tensor, transposed: bool = False, n_split: int = 1, _flip_y: bool = True,
                        _planes=None) -> torch.Tensor:
    """torch conv weight -> bf16 [sel][Cin/16][j][plane hi/lo][tz][khalf][ty'][n_cta][8], the shared-memory
    image vsseg_conv3d_tc streams with cp.async.bulk (layout documented in include/vsseg_b200.h).
    ty' = ky-1-ty for Conv3d (so the y taps that share an input line are adjacent N rows), ty for
    ConvTranspose3d."""
    if transposed:  # ConvTranspose3d [Cin, Cout, kx, ky, kz] -> per output-x parity px, input-x shift j
        wt = w.permute(1, 0, 2, 3, 4)
        z = torch.zeros_like(wt[:, :, 0])
        per_px = [torch.stack([wt[:, :, 1], z], 2), torch.stack([wt[:, :, 2], wt[:, :, 0]], 2)]
        return torch.cat([pack_conv_weight_tc(q, False, n_split, _flip_y=False) for q in per_px]).contiguous()
    cout, cin, kx, ky, kz = w.shape
    cp = _round_up(cout, 16)
    if cp % n_split or (cp // n_split) % 16 or cin % 16:
        raise ValueError("pack_conv_weight_tc: Cin % 16, round_up(Cout,16) % (16*n_split) must be 0")
    if cp != cout:
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, 0, 0, cp - cout))
    n_cta = cp // n_split
    if _flip_y:
        w = w.flip(3)
    planes = _split_planes(w) if _planes is None else _planes(w)
    p = planes.reshape(2, n_split, n_cta, cin // 16, 2, 8, kx, ky, kz)  # plane sel n c khalf e j ty tz
    return p.permute(1, 3, 6, 0, 8, 4, 7, 2, 5).contiguous()  # sel c j plane tz khalf ty' n e


def pack_conv_weight_tc_2p(w: torch.Tensor) -> torch.Tensor:
    """Two-pass image for vsseg_conv3d_tc_f32out_2p (Cout <= 2): plane 0 columns = [hi(W) | lo(W) | 0],
    plane 1 columns = [hi(W) | 0]; both planes hold exact bf16 values."""
    c = w.shape[0]
    if c > 8:
        raise ValueError("two-pass packing needs Cout <= 8")

    def planes(wp):   # wp: [16, Cin, kx, ky, kz] (zero-padded, y-flipped), rows 0..c-1 real
        hl = _split_planes(wp[:c])
        p0 = torch.zeros_like(wp, dtype=torch.bfloat16)
        p1 = torch.zeros_like(p0)
        p0[:c], p0[c:2 * c], p1[:c] = hl[0], hl[1], hl[0]
        return torch.stack([p0, p1])

    return pack_conv_weight_tc(w, False, 1, _planes=planes)


def pack_shortcut_weight_tc(w: torch.Tensor, n_split: int = 1) -> torch.Tensor:
    """1x1x1 shortcut Conv3d weight [Cout, Csrc, 1,1,1] -> bf16 [n-slice][Csrc/16][plane][khalf][n_cta][8]."""
    return pack_conv_weight_tc(w, False, n_split).reshape(n_split, w.shape[1] // 16, 2, 2, -1, 8).contiguous()


def tc_enabled() -> bool:
    import os
    return os.environ.get("VSSEG_TC", "1") != "0"


def fold_epilogue(sd, p, cout, cout_pad, norm: bool, act: str):
    """Per-channel scale/shift of conv bias + eval BatchNorm3d (eps 1e-5), and the activation."""
    bias = sd[p + "conv.bias"].double()
    if norm:
        g, b = sd[p + "norm.weight"].double(), sd[p + "norm.bias"].double()
        m, v = sd[p + "norm.running_mean"].double(), sd[p + "norm.running_var"].double()
        scale = g / torch.sqrt(v + 1e-5)
        shift = b + (bias - m) * scale
    else:
        scale = torch.ones(cout, dtype=torch.float64, device=bias.device)
        shift = bias
    pad = (0, cout_pad - cout)
    scale = torch.nn.functional.pad(scale.float(), pad, value=1.0)
    shift = torch.nn.functional.pad(shift.float(), pad)
    if act == "prelu":
        slope, code = float(sd[p + "act.weight"].reshape(-1)[0]), 0
    elif act == "relu":
        slope, code = 0.0, 0
    elif act == "none":
        slope, code = 1.0, 0
    elif act == "sigmoid":
        slope, code = 0.0, 1
    else:
        raise ValueError(act)
    return scale.contiguous(), shift.contiguous(), slope, code


class _Step:
    __slots__ = ("fn", "args", "name", "flops", "bytes", "kind")

    def __init__(self, name, fn, args, flops=0, nbytes=0, kind="generic"):
        self.name, self.fn, self.args = name, fn, args
        self.flops, self.bytes, self.kind = int(flops), int(nbytes), kind


def _nvox(v):
    return v.B * v.X * v.Y * v.Z


def _conv_cost(src, dst, k, transposed, cin, cout, extra_elems=0):
    """Algorithmic cost of one conv launch: 2*MACs, and bytes = every activation element read once
    and written once at 4 B (split-bf16 or fp32) plus fp32 weights."""
    taps = k[0] * k[1] * k[2]
    macs = (_nvox(src) if transposed else _nvox(dst)) * taps * cin * cout
    nbytes = 4 * (_nvox(src) * cin + _nvox(dst) * cout + extra_elems + taps * cin * cout)
    return 2 * macs, nbytes


def batch_first_enabled():
    """VSSEG_SW_BATCH_FIRST: window-group plans run the first ResidualUnit of all windows in one launch per conv
    (window set of source views) instead of one launch per window."""
    return os.environ.get("VSSEG_SW_BATCH_FIRST", "1") == "1"


class UNetEvalPlan:
    """Flat launch list for one patch shape [B,1,X,Y,Z] (X,Y % 32 == 0, Z % 8 == 0).

    ``window_levels`` = d > 0 makes the B entries independent WINDOWS (sliding-window groups): the d
    finest levels are launched per window on its own source/destination view (each such launch already
    fills the GPU), the coarser levels once for all B windows - their launches are latency-bound on a
    single window (a few dozen CTAs, K walked serially), so B windows cost about the same as one.
    """

    def __init__(self, state_dict, patch_size, batch=1, device="cuda:0", attention=True, window_levels=0,
                 channels=(16, 32, 48, 64, 80, 96),
                 strides=((2, 2, 1), (2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2)),
                 kernel_sizes=((3, 3, 1), (3, 3, 1), (3, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                 sample_kernel_sizes=((3, 3, 1), (3, 3, 1), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                 in_channels=1, out_channels=2, tensor_cores=None, atomic_out=False):
        self.lib = _lib.load()
        self.use_tc = tc_enabled() if tensor_cores is None else bool(tensor_cores)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NativeLibraryError("UNetEvalPlan runs on a CUDA device only (no CPU fallback)")
        if in_channels != 1 or out_channels not in (1, 2):
            raise NotImplementedError("native plan supports in_channels=1, out_channels in {1,2}")
        if channels[0] != 16 or any(c % 8 for c in channels):
            raise NotImplementedError("native plan needs channels[0]==16 and channels % 8 == 0")
        self.attention = bool(attention)
        self.channels, self.strides = tuple(channels), tuple(tuple(s) for s in strides)
        self.kernel_sizes = tuple(tuple(k) for k in kernel_sizes)
        self.sample_kernel_sizes = tuple(tuple(k) for k in sample_kernel_sizes)
        self.out_channels = out_channels
        self.B = int(batch)
        self.patch = tuple(int(v) for v in patch_size)
        tot = [math.prod(s[d] for s in self.strides) for d in range(3)]
        if any(self.patch[d] % tot[d] for d in range(3)):
            raise ValueError(f"patch {self.patch} must be divisible by {tuple(tot)} "
                             "(the reference's skip torch.cat fails otherwise)")
        self.sd = {k: v.detach().to(self.device) for k, v in state_dict.items()}
        self._keep = []  # device tensors referenced by raw pointers
        self.steps = []
        self.att_maps = []  # fp32 [B,1,x,y,z] tensors, coarsest first (hook order)
        # the per-call descriptors (mutated in place by run()): one pair for the whole batch, or one pair
        # per window when the fine levels run per window
        self.window_levels = max(0, min(int(window_levels), len(channels) - 1)) if self.B > 1 else 0
        nview = self.B if self.window_levels else 1
        # contiguous records: one launch can take every window's source (a window set, vsseg_f32view.n_windows) /
        # destination
        self._src_arr = (_lib.F32View * nview)()
        self.srcs = [self._src_arr[i] for i in range(nview)]
        self._src_set = False   # the first ResidualUnit reads all windows in one launch (set by _build)
        self._dst_arr = (_lib.F32View * nview)()
        self.dsts = [self._dst_arr[i] for i in range(nview)]
        self.src, self.dst = self.srcs[0], self.dsts[0]
        self._bs = (0, self.B)   # batch slice the step being emitted works on
        self.sw_weight = C.c_void_p(None)
        self.atomic_blend = C.c_int32(0)   # set per run(): blend with atomics (several writers of one accumulator)
        self._has_plain_blend = False      # a blending launch that cannot use atomics is part of the plan
        # the caller always blends with atomics (multi-GPU peer accumulator): the gate+logits launch then takes every
        # window of the group at once (2048 CTAs instead of 8 x 256 on 296 slots: 0.71 -> 0.55 ms per group of 8)
        self.atomic_out = bool(atomic_out)
        self._build()

    # -- helpers ------------------------------------------------------------------------
    def _dev(self, t):
        t = t.to(self.device).contiguous()
        self._keep.append(t)
        return t

    def _geom(self, k, s=(1, 1, 1), transposed=False):
        return _lib.ConvGeom(k[0], k[1], k[2], s[0], s[1], s[2], 1 if transposed else 0)

    def _buf(self, C_, dims):
        b = Act8Buffer(self.B, C_, dims[0], dims[1], dims[2], self.device)
        self._keep.append(b)
        return b

    def _v(self, buf, c0=0, C_=None):
        """View of the batch slice the current step works on."""
        return buf.view(c0, C_, self._bs[0], self._bs[1])

    def _add_conv(self, name, p, src, dst, k, stride=(1, 1, 1), transposed=False, norm=True, act="prelu",
                  res=None, res_cin1=None, shortcut=None):
        """src/dst: Act8 views.  One fused Convolution block (+ optional residual).
        shortcut = (prefix of the 1x1x1 residual conv, its Act8 source): fused as a second accumulator
        on the tensor-core path; returns False if it could not be fused (caller adds it separately)."""
        wshape = self.sd[p + "conv.weight"].shape
        cout = wshape[1] if transposed else wshape[0]  # dst may carry zero-padded extra channels (cout < dst.C)
        cpad = _round_up(dst.C, 16)
        scale, shift, slope, code = fold_epilogue(self.sd, p, cout, cpad, norm, act)
        scale, shift = self._dev(scale), self._dev(shift)
        ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), slope, code)
        g = self._geom(k, stride, transposed)
        res_p = C.byref(res) if res is not None else None
        tail = (None, C.byref(self._cur_src()), res_cin1[0].data_ptr(), res_cin1[1].data_ptr()) if res_cin1 is not None \
            else (res_p, None, None, None)
        self._keep += [src, dst, g, ep, res]
        extra = _nvox(dst) * cout if res is not None else (_nvox(dst) if res_cin1 is not None else 0)
        fl, nb = _conv_cost(src, dst, k, transposed, src.C, cout, extra)
        if self.use_tc and src.C % 16 == 0:
            sc_src = shortcut[1] if shortcut is not None else None
            sc_p = C.byref(sc_src) if sc_src is not None else None
            ns = self.lib.vsseg_conv3d_tc_suggest_split(C.byref(src), C.byref(dst), C.byref(g), sc_p)
            fused = ns > 0 and shortcut is not None
            if ns == 0 and shortcut is not None:
                sc_p = None
                ns = self.lib.vsseg_conv3d_tc_suggest_split(C.byref(src), C.byref(dst), C.byref(g), None)
            if ns > 0:
                w = self._dev(pack_conv_weight_tc(self._pad_cout(self.sd[p + "conv.weight"], transposed, dst.C),
                                                  transposed, ns))
                sc_tail = (None, None, None)
                if fused:
                    q = shortcut[0]
                    w2 = self._dev(pack_shortcut_weight_tc(self.sd[q + "weight"], ns))
                    b2 = self._dev(torch.nn.functional.pad(self.sd[q + "bias"].float(), (0, cpad - cout)))
                    sc_tail = (sc_p, w2.data_ptr(), b2.data_ptr())
                    self._keep.append(sc_src)
                    f2, n2 = _conv_cost(sc_src, dst, (1, 1, 1), False, sc_src.C, cout)
                    # the output is written once; a decoder unit's shortcut reads the SAME tensor as its conv
                    # (one subunit), which SURVEY.md §8d counts once
                    same = sc_src.hi == src.hi and sc_src.C == src.C
                    fl, nb = fl + f2, nb + n2 - 4 * _nvox(dst) * cout - (4 * _nvox(sc_src) * sc_src.C if same else 0)
                args = (C.byref(src), C.byref(dst), C.byref(g), w.data_ptr(), ns, C.byref(ep)) + tail + sc_tail
                self.steps.append(_Step(name, self.lib.vsseg_conv3d_tc, args, fl, nb, kind="tcgen05"))
                return fused
        w = self._dev(pack_conv_weight(self.sd[p + "conv.weight"], transposed, cpad))
        args = (C.byref(src), C.byref(dst), C.byref(g), w.data_ptr(), cpad, C.byref(ep)) + tail
        self.steps.append(_Step(name, self.lib.vsseg_conv3d_act8, args, fl, nb))
        return False

    def _tag(self):
        return f"@w{self._bs[0]}" if self.window_levels and self._bs[1] == 1 else ""

    def _cur_src(self):
        return self.srcs[self._bs[0]] if self.window_levels and self._bs[1] == 1 else self.srcs[0]

    def _cur_dst(self):
        return self.dsts[self._bs[0]] if self.window_levels and self._bs[1] == 1 else self.dsts[0]

    @staticmethod
    def _pad_cout(w, transposed, c):
        """Zero-pad the output channels of a conv weight to c (extra channels of the act8 buffer stay 0)."""
        d = 1 if transposed else 0
        if w.shape[d] == c:
            return w
        pad = [0, 0] * (w.dim() - 1 - d) + [0, c - w.shape[d]]
        return torch.nn.functional.pad(w, pad)

    def _add_smallcout(self, name, src, out_view, k, w, bias, act_code, slope, sw_weight=None, nb_extra=0, gate=None):
        """Conv with 1-2 output channels -> planar fp32: tensor-core path when the shape is covered
        (weights zero-padded to Cin = src.C, Cout = 16), else the CUDA-core kernel.
        gate: act8 view that AttentionBlock2 scales by 1 + out (Cout = 1); returns True when the gate was fused
        into this launch (tensor-core path), else the caller adds the gate launch."""
        g = self._geom(k)
        cout = w.shape[0]
        fl, nb = _conv_cost(src, src, k, False, src.C, cout)  # stride 1: output extents = input extents
        nb += nb_extra
        if w.shape[1] != src.C:  # hidden buffer padded to a multiple of 16 channels
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, src.C - w.shape[1]))
        self._keep += [src, out_view, g]
        if self.use_tc and src.C % 16 == 0:
            o16 = _lib.Act8(src.hi, 0, 0, src.B, 16, src.X, src.Y, src.Z)  # extents only (plan check)
            if self.lib.vsseg_conv3d_tc_supported(C.byref(src), C.byref(o16), C.byref(g), 1, None):
                wp = self._dev(pack_conv_weight_tc_2p(w))
                scale = self._dev(torch.ones(16))
                shift = self._dev(torch.nn.functional.pad(bias.float(), (0, 16 - cout)))
                ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), slope, act_code)
                self._keep.append(ep)
                # measured (group of 8 windows): with the loads of four channel groups in flight ahead of the first
                # store the fused gate wins down to 64x64x128 (conv2 + gate 0.160 + 0.383 ms -> 0.488 ms fused; level 3:
                # -9 us per window); at 128^3 the gate is applied on the fly by the gate+logits launch, and when that is
                # off it runs faster as its own bandwidth-bound launch
                fuse_env = os.environ.get("VSSEG_FUSE_GATE", "auto")
                fuse = fuse_env == "1" or (fuse_env == "auto" and gate is not None and _nvox(gate) // gate.B <= 64 * 64 * 128)
                if gate is not None and cout == 1 and sw_weight is None and fuse:
                    self._keep.append(gate)
                    self.steps.append(_Step(name + "+gate", self.lib.vsseg_conv3d_tc_attgate,
                                            (C.byref(src), C.byref(out_view), C.byref(g), wp.data_ptr(), C.byref(ep),
                                             C.byref(gate)), fl + 2 * _nvox(gate) * gate.C,
                                            nb + 8 * _nvox(gate) * gate.C, kind="tcgen05"))
                    return True
                self.steps.append(_Step(name, self.lib.vsseg_conv3d_tc_f32out_2p,
                                        (C.byref(src), C.byref(out_view), C.byref(g), wp.data_ptr(), C.byref(ep),
                                         sw_weight), fl, nb, kind="tcgen05"))
                return False
        wg = self._dev(pack_conv_weight(w, False))
        b = self._dev(bias.float())
        self.steps.append(_Step(name, self.lib.vsseg_conv3d_smallcout,
                                (C.byref(src), C.byref(out_view), C.byref(g), wg.data_ptr(), b.data_ptr(), act_code,
                                 slope, sw_weight), fl, nb))
        return False

    def _batch_first_ok(self, hbuf, catbuf, c0):
        if not batch_first_enabled() or self.B > _lib.MAX_WINDOWS or not self.use_tc:
            return False
        # the Cin=1 shortcut of unit1 takes a window set on the tensor-core path only
        g = self._geom(self.kernel_sizes[0])
        h, e = hbuf.view(), catbuf.view(0, c0)
        return bool(self.lib.vsseg_conv3d_tc_supported(C.byref(h), C.byref(e), C.byref(g), 1, None))

    def _gate_logits_ok(self, buf, k):
        return (os.environ.get("VSSEG_GATE_LOGITS", "1") != "0" and buf.C == 32 and tuple(k) == (3, 3, 1)
                and buf.Z % 8 == 0)

    def _add_gate_logits(self, name, src, k, w, bias):
        """Last ResidualUnit (conv_only conv + folded shortcut) with the attention gate applied on the fly and the
        sliding-window blend (vsseg_conv3d_gate_logits).  One launch covers every window of the batch slice."""
        cout = w.shape[0]
        wh = pack_conv_weight(w, False).cpu().contiguous()          # HOST fp32 [9][Cin][Cout], passed by value
        bh = bias.detach().float().cpu().contiguous()
        att_p = None
        if self.attention:
            att = self._att["dec0.att"]
            av = f32view(att[self._bs[0]:self._bs[0] + self._bs[1]])
            att_p = C.byref(av)
            self._keep.append(av)
        whole = self._bs == (0, self.B)
        atomic = self.atomic_blend
        if self.window_levels and whole:
            # every window of the group in ONE launch: their destination regions overlap, so the blend must be
            # atomic (sum order unspecified).  Opt-in: the default keeps one launch per window = MONAI's order
            outs, n_outs, atomic = self._dst_arr, self.B, 1
        else:
            outs, n_outs = C.byref(self._cur_dst()), 1
        self._keep += [wh, bh, src]
        nv = _nvox(src)
        fl = 2 * nv * 9 * src.C * cout + (2 * nv * src.C if att_p is not None else 0)
        nb = 4 * (nv * (src.C + (1 if att_p is not None else 0)) + nv * (2 * cout + 1) + 9 * src.C * cout)
        self.steps.append(_Step(name, self.lib.vsseg_conv3d_gate_logits,
                                (C.byref(src), att_p, wh.data_ptr(), bh.data_ptr(), cout, outs, n_outs, self.sw_weight, atomic),
                                fl, nb))

    def _add_shortcut(self, name, p, src, dst):
        """1x1x1 shortcut conv of a ResidualUnit (convolutions.py:241-250) -> addend buffer."""
        cout = dst.C
        cpad = _round_up(cout, 16)
        w = self._dev(pack_conv_weight(self.sd[p + "weight"], False, cpad))
        scale = self._dev(torch.ones(cpad))
        shift = self._dev(torch.nn.functional.pad(self.sd[p + "bias"].float(), (0, cpad - cout)))
        ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), 1.0, 0)
        g = self._geom((1, 1, 1))
        args = (C.byref(src), C.byref(dst), C.byref(g), w.data_ptr(), cpad, C.byref(ep), None, None, None, None)
        self._keep += [src, dst, g, ep]
        fl, nb = _conv_cost(src, dst, (1, 1, 1), False, src.C, cout)
        self.steps.append(_Step(name, self.lib.vsseg_conv3d_act8, args, fl, nb))

    def _add_att(self, name, p, buf, hid, k, gate=True):
        """AttentionBlock1+2 on act8 buffer `buf` (all channels), in place.  The hidden tensor uses a
        multiple of 16 channels (zero weights for the padding) so both convs run on tensor cores.
        gate=False: only the map is computed; the consumer applies x*(1+att) on the fly (top level)."""
        cin = buf.C
        src, h = self._v(buf), self._v(hid, 0, _round_up(cin // 2, 16))
        self._add_conv(name + ".conv1", p + "0.conv1.", src, h, k, norm=False, act="relu")
        key = name.split("@")[0]   # one map per gate, shared by the per-window steps
        att = self._att.get(key)
        if att is None:
            att = self._att[key] = torch.empty((self.B, 1, buf.X, buf.Y, buf.Z), dtype=torch.float32, device=self.device)
            self.att_maps.append(att)
        av = f32view(att[self._bs[0]:self._bs[0] + self._bs[1]])
        fused = self._add_smallcout(name + ".conv2", h, av, k, self.sd[p + "0.conv2.conv.weight"],
                                    self.sd[p + "0.conv2.conv.bias"], 1, 0.0, gate=src if gate else None)
        if gate and not fused:
            self.steps.append(_Step(name + ".gate", self.lib.vsseg_att_gate, (C.byref(src), C.byref(av), C.byref(src)),
                                    2 * _nvox(src) * cin, 4 * _nvox(src) * (2 * cin + 1)))
        self._keep += [src, h, av]

    def _add_ru(self, name, p, src, h_buf, r_buf, dst, k, subunits):
        """ResidualUnit with `subunits` Convolution blocks and a 1x1x1 shortcut; src/dst act8 views.
        The shortcut is a second accumulator of the last conv when the tensor-core path covers it,
        else a separate launch into the addend buffer r_buf."""
        cout = dst.C
        sc = (p + "residual.", src)
        last = p + f"conv.unit{subunits - 1}."
        last_src = src if subunits == 1 else self._v(h_buf, 0, cout)
        if subunits == 2:
            self._add_conv(name + ".unit0", p + "conv.unit0.", src, last_src, k)
        n0 = len(self.steps)
        if self._add_conv(name + f".unit{subunits - 1}", last, last_src, dst, k, shortcut=sc):
            return
        # not fused: drop the step just added and redo it with an explicit shortcut launch
        del self.steps[n0:]
        r = self._v(r_buf, 0, cout)
        self._add_shortcut(name + ".residual", p + "residual.", src, r)
        self._add_conv(name + f".unit{subunits - 1}", last, last_src, dst, k, res=r)

    # -- plan ---------------------------------------------------------------------------
    def _build(self):
        ch, nlev = self.channels, len(self.channels)
        dims = [self.patch]
        for s in self.strides:
            dims.append(tuple(d // q for d, q in zip(dims[-1], s)))
        cat = [self._buf(2 * ch[l], dims[l]) for l in range(nlev - 1)]
        hb = [self._buf(ch[l], dims[l]) for l in range(nlev - 1)]
        rb = [self._buf(ch[l], dims[l]) for l in range(nlev - 1)]
        dn = [self._buf(ch[l], dims[l + 1]) for l in range(nlev - 1)]
        bot_h = self._buf(ch[-1], dims[-1])
        bot_r = self._buf(ch[-1], dims[-1])
        bot_o = self._buf(ch[-1], dims[-1])
        bot_hid = self._buf(_round_up(ch[-2] // 2, 16), dims[-1])
        self.buffers = dict(cat=cat, h=hb, r=rb, down=dn, bot_h=bot_h, bot_r=bot_r, bot_o=bot_o)

        prefixes = []
        p = "model."
        for l in range(nlev - 1):
            prefixes.append(p)
            p = p + "1.submodule.1."
        self._att = {}
        v = self._v

        def encoder(l, part="all"):
            p, k, sk, s = prefixes[l], self.kernel_sizes[l], self.sample_kernel_sizes[l], self.strides[l]
            tag = self._tag()
            e = v(cat[l], 0, ch[l])
            if part == "down":
                self._add_conv(f"down{l}" + tag, p + "1.submodule.0.", e, v(dn[l]), sk, stride=s)
                return
            if l == 0:
                # unit0 reads the 1-channel fp32 source in place; the 1x1x1 shortcut (Cin=1) is an
                # affine map of the same source applied in unit1's epilogue.
                q = p + "0.conv.unit0."
                w = self._dev(pack_conv_weight(self.sd[q + "conv.weight"], False)[:, 0, :].contiguous())
                scale, shift, slope, code = fold_epilogue(self.sd, q, ch[0], ch[0], True, "prelu")
                scale, shift = self._dev(scale), self._dev(shift)
                ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), slope, code)
                g = self._geom(k)
                h = v(hb[0])
                self._keep += [ep, g, h]
                fl, nb = _conv_cost(h, h, k, False, 1, ch[0])
                self.steps.append(_Step("enc0.unit0" + tag, self.lib.vsseg_conv3d_cin1,
                                        (C.byref(self._cur_src()), C.byref(h), C.byref(g), w.data_ptr(), C.byref(ep)),
                                        fl, nb))   # 4 B read + 64 B written per voxel
                rw = self._dev(self.sd[p + "0.residual.weight"].reshape(-1).float())
                rbias = self._dev(self.sd[p + "0.residual.bias"].float())
                self._add_conv("enc0.unit1" + tag, p + "0.conv.unit1.", h, e, k, res_cin1=(rw, rbias))
            else:
                self._add_ru(f"enc{l}" + tag, p + "0.", v(dn[l - 1]), hb[l], rb[l], e, k, 2)
            if part != "units":
                self._add_conv(f"down{l}" + tag, p + "1.submodule.0.", e, v(dn[l]), sk, stride=s)

        def decoder(l, part="all"):
            p, k, sk, s = prefixes[l], self.kernel_sizes[l], self.sample_kernel_sizes[l], self.strides[l]
            tag = self._tag()
            sub_out = v(hb[l + 1], 0, ch[l + 1]) if l + 1 < nlev - 1 else v(bot_o)
            if part != "out":
                self._add_conv(f"up{l}" + tag, p + "1.submodule.2.", sub_out, v(cat[l], ch[l], ch[l]), sk, stride=s,
                               transposed=True)
            pr = p + "2."
            # top level: the gated tensor feeds only the logits conv, so the gate is applied on the fly by
            # vsseg_conv3d_gate_logits and never written (no dec0.att.gate launch)
            fuse_top = l == 0 and self._gate_logits_ok(cat[0], k)
            if self.attention:
                if part != "out":
                    self._add_att(f"dec{l}.att" + tag, pr + "0.", cat[l], hb[l], k, gate=not fuse_top)
                pr = pr + "1."
            if part == "upatt":
                return
            if l > 0:
                self._add_ru(f"dec{l}" + tag, pr, v(cat[l]), None, rb[l], v(hb[l], 0, ch[l]), k, 1)
            else:
                # top unit: conv_only + shortcut, both linear -> one conv (shortcut folded into the
                # centre tap), written to planar fp32 or blended into the sliding-window accumulator.
                w = self.sd[pr + "conv.unit0.conv.weight"].float().clone()
                w[:, :, k[0] // 2, k[1] // 2, k[2] // 2] += self.sd[pr + "residual.weight"].reshape(
                    self.out_channels, -1).float()
                bias = self.sd[pr + "conv.unit0.conv.bias"] + self.sd[pr + "residual.bias"]
                src = v(cat[0])
                if fuse_top:
                    self._add_gate_logits("dec0.gate+logits" + tag, src, k, w, bias)
                    return
                # bytes: out read-modify-write + weight map instead of a plain store
                self._has_plain_blend = True
                self._add_smallcout("dec0.logits" + tag, src, self._cur_dst(), k, w, bias, 0, 1.0, self.sw_weight,
                                    nb_extra=4 * _nvox(src) * (self.out_channels + 1))

        d = self.window_levels
        windows = [(i, 1) for i in range(self.B)] if d else []
        # with d = 1 only the launches that touch a window's own source / destination view run per window
        # (first ResidualUnit, logits conv); the rest of the finest level is batched like the coarse ones
        split0 = d == 1
        # ---- fine levels of the encoder, window by window
        # VSSEG_SW_BATCH_FIRST=1: the first ResidualUnit of every window in ONE launch per conv - the windows' source
        # views travel as a window set (include/vsseg_b200.h).  A single 128^3 window is 128-144 tiles for 148
        # persistent CTAs (one tile each: the pipeline ramp is never amortised); the group is 1024 tiles.
        self._src_set = bool(split0 and self._batch_first_ok(hb[0], cat[0], ch[0]))
        if self._src_set:
            self._bs = (0, self.B)
            encoder(0, "units")
        else:
            for bs in windows:
                self._bs = bs
                for l in range(d):
                    encoder(l, "units" if split0 else "all")
        # ---- coarse levels, all windows at once
        self._bs = (0, self.B)
        if split0:
            encoder(0, "down")
        for l in range(d, nlev - 1):
            encoder(l)
        pb = prefixes[-1] + "1.submodule.1."
        kb = self.kernel_sizes[-1]
        if self.attention:
            self._add_att("bottom.att", pb + "0.", dn[-1], bot_hid, kb)
            self._add_ru("bottom", pb + "1.", v(dn[-1]), bot_h, bot_r, v(bot_o), kb, 2)
        else:
            self._add_ru("bottom", pb, v(dn[-1]), bot_h, bot_r, v(bot_o), kb, 2)
        for l in range(nlev - 2, d - 1, -1):
            decoder(l)
        if split0:
            decoder(0, "upatt")
        # ---- fine levels of the decoder, window by window (VSSEG_SW_ATOMIC=1: the fused gate+logits launch takes all
        # windows at once and blends with atomics)
        if (split0 and self.B <= 16 and self._gate_logits_ok(cat[0], self.kernel_sizes[0])
                and (self.atomic_out or os.environ.get("VSSEG_SW_ATOMIC", "0") == "1")):
            decoder(0, "out")
            windows = []
        for bs in windows:
            self._bs = bs
            for l in range(d - 1, -1, -1):
                decoder(l, "out" if split0 else "all")
        self._bs = (0, self.B)
        del self.sd

    # -- execution ----------------------------------------------------------------------
    def _set(self, view, new):
        C.memmove(C.byref(view), C.byref(new), C.sizeof(_lib.F32View))

    def _bind(self, src, dst, sw_weight_ptr, atomic=False):
        srcs = list(src) if isinstance(src, (list, tuple)) else [src]
        dsts = list(dst) if isinstance(dst, (list, tuple)) else [dst]
        nb = 1 if self.window_levels else self.B
        if len(srcs) != len(self.srcs) or len(dsts) != len(self.dsts):
            raise ValueError(f"run(): the plan takes {len(self.srcs)} source/destination view(s)")
        for s_, d_ in zip(srcs, dsts):
            if (s_.B, s_.X, s_.Y, s_.Z) != (nb, *self.patch) or (d_.B, d_.C, d_.X, d_.Y, d_.Z) != (
                    nb, self.out_channels, *self.patch):
                raise ValueError("run(): view shapes do not match the plan")
        for mine, new in zip(self.srcs + self.dsts, srcs + dsts):
            self._set(mine, new)
        if self._src_set:   # record 0 heads the window set of all B sources
            if any(s_.n_windows > 1 for s_ in srcs):
                raise ValueError("run(): source views must be plain single-window views")
            self._src_arr[0].n_windows = self.B
        self.sw_weight.value = sw_weight_ptr
        if atomic and self._has_plain_blend:
            raise _lib.NativeLibraryError("this plan blends with a launch that has no atomic mode")
        if self.atomic_out and not atomic and sw_weight_ptr:
            raise ValueError("run(): a plan built with atomic_out blends with atomics only")
        self.atomic_blend.value = 1 if atomic else 0

    def run(self, src, dst, sw_weight_ptr: int | None = None, stream=None, count=True, atomic=False):
        """Launch the whole forward for one patch batch.

        src: [B,1,X,Y,Z] fp32 region (may be a strided window of a larger volume);
        dst: [B,out_channels,X,Y,Z] fp32 region; if ``sw_weight_ptr`` is given the logits are
        blended (dst += weight * logits) instead of stored.  A plan built with ``window_levels`` takes
        lists of B single-window views instead.  ``atomic``: blend with red.global.add (several ranks write one
        accumulator over peer memory).
        """
        self._bind(src, dst, sw_weight_ptr, atomic)
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        for st in self.steps:
            code = st.fn(*st.args, s)
            if code:
                _lib.check(code, st.name)
        if count:
            _lib.count_launch(len(self.steps))

    def profile(self, src, dst, sw_weight_ptr=None, iters=3):
        """Per-step device time (ms, mean of `iters`, CUDA events on the launch stream)."""
        self._bind(src, dst, sw_weight_ptr)
        stream = torch.cuda.current_stream(self.device)
        s = stream.cuda_stream
        ms = [0.0] * len(self.steps)
        for it in range(iters + 1):
            # an untimed pass is queued right before the timed one so the GPU is busy while the timed launches are
            # issued: otherwise the first step's interval would include the host's launch latency
            for st in self.steps:
                _lib.check(st.fn(*st.args, s), st.name)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.steps) + 1)]
            evs[0].record(stream)
            for i, st in enumerate(self.steps):
                _lib.check(st.fn(*st.args, s), st.name)
                evs[i + 1].record(stream)
            torch.cuda.synchronize(self.device)
            if it:  # first pass is a warm-up
                for i in range(len(self.steps)):
                    ms[i] += evs[i].elapsed_time(evs[i + 1]) / iters
        return [(st.name, st.kind, st.flops, st.bytes, t) for st, t in zip(self.steps, ms)]

    def total_flops(self):
        return sum(st.flops for st in self.steps)

    def forward(self, x: torch.Tensor):
        """x: [B,1,X,Y,Z] fp32 CUDA tensor -> (logits [B,out,X,Y,Z], att_maps coarsest first)."""
        if x.device != self.device or x.dtype != torch.float32:
            raise ValueError("forward(): expects a float32 tensor on the plan's device")
        out = torch.empty((self.B, self.out_channels, *self.patch), dtype=torch.float32, device=self.device)
        if self.window_levels:
            self.run([f32view(x[i:i + 1]) for i in range(self.B)], [f32view(out[i:i + 1]) for i in range(self.B)])
        else:
            self.run(f32view(x), f32view(out))
        return out, list(self.att_maps)


def conv_block_ncdhw(x, sd, kernel_size, stride, transposed, norm, act, residual=None, shortcut=None,
                     require_tc=False):
    """One fused Convolution block on an NCDHW fp32 CUDA tensor (standalone-module path).

    sd holds conv.weight / conv.bias (/ norm.* / act.weight).  Channels are zero-padded to the
    act8 granularity (8 in, 16 out) so any channel count works; pack and unpack are native kernels.
    shortcut = (x_src NCDHW, weight [Cout,Csrc,1,1,1], bias [Cout]) fuses a 1x1x1 shortcut conv as a
    second accumulator (tensor-core path only); x_src may be `x` itself (decoder units).  require_tc raises if the tcgen05 path does not cover
    the shape (used by the tests to prove which kernel ran).
    """
    lib = _lib.load()
    if not x.is_cuda:
        raise _lib.NativeLibraryError("conv_block_ncdhw needs a CUDA tensor (no CPU fallback)")
    w0 = sd["conv.weight"]
    cout = w0.shape[1] if transposed else w0.shape[0]
    B, cin = x.shape[0], x.shape[1]
    cin8, cout16 = _round_up(cin, 8), _round_up(cout, 16)
    sd = dict(sd)
    if sd.get("conv.bias") is None:
        sd["conv.bias"] = torch.zeros(cout, device=x.device)
    w = pack_conv_weight(w0, transposed, cout16)
    if cin8 != cin:
        w = torch.nn.functional.pad(w, (0, 0, 0, cin8 - cin))
    w = w.contiguous()
    scale, shift, slope, code = fold_epilogue(sd, "", cout, cout16, norm, act)
    xin = x.float()
    if cin8 != cin:
        xin = torch.nn.functional.pad(xin, (0, 0, 0, 0, 0, 0, 0, cin8 - cin))
    src = Act8Buffer(B, cin8, *x.shape[2:], x.device).from_ncdhw(xin)
    if transposed:
        odims = [d * s for d, s in zip(x.shape[2:], stride)]
    else:
        odims = [(d + s - 1) // s for d, s in zip(x.shape[2:], stride)]
    dst = Act8Buffer(B, cout16, *odims, x.device)
    res_v = None
    if residual is not None:
        r = torch.nn.functional.pad(residual.float(), (0, 0, 0, 0, 0, 0, 0, cout16 - cout))
        rbuf = Act8Buffer(B, cout16, *odims, x.device).from_ncdhw(r)
        res_v = rbuf.view()
    g = _lib.ConvGeom(*[int(k) for k in kernel_size], *[int(s) for s in stride], 1 if transposed else 0)
    ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), slope, code)
    sv, dv = src.view(), dst.view()
    stream = torch.cuda.current_stream(x.device).cuda_stream
    res_p = C.byref(res_v) if res_v is not None else None
    ns = 0
    sc_v = None
    if shortcut is not None:
        xs, ws, bs = shortcut
        if xs is x:   # a decoder unit: the shortcut reads the conv's own input (no shortcut stages in the kernel)
            sc_v = src.view()
        else:
            sbuf = Act8Buffer(B, xs.shape[1], *xs.shape[2:], x.device).from_ncdhw(xs.float())
            sc_v = sbuf.view()
    if tc_enabled() and cin8 == cin and cin % 16 == 0:
        ns = lib.vsseg_conv3d_tc_suggest_split(C.byref(sv), C.byref(dv), C.byref(g),
                                               C.byref(sc_v) if sc_v is not None else None)
    if (require_tc or shortcut is not None) and ns <= 0:
        raise _lib.NativeLibraryError("the tcgen05 path does not cover this shape")
    if ns > 0:
        wt = pack_conv_weight_tc(w0, transposed, ns)
        sc_tail = (None, None, None)
        if shortcut is not None:
            w2 = pack_shortcut_weight_tc(ws, ns)
            b2 = torch.nn.functional.pad(bs.float(), (0, cout16 - cout)).contiguous()
            sc_tail = (C.byref(sc_v), w2.data_ptr(), b2.data_ptr())
        _lib.check(lib.vsseg_conv3d_tc(C.byref(sv), C.byref(dv), C.byref(g), wt.data_ptr(), ns, C.byref(ep), res_p,
                                       None, None, None, *sc_tail, stream), "conv3d_tc")
    else:
        _lib.check(lib.vsseg_conv3d_act8(C.byref(sv), C.byref(dv), C.byref(g), w.data_ptr(), cout16, C.byref(ep),
                                         res_p, None, None, None, stream), "conv3d_act8")
    _lib.count_launch()
    return dst.to_ncdhw(0, cout16)[:, :cout].contiguous()


def native_shortcut(conv: torch.nn.Module, x):
    """The ResidualUnit shortcut conv (reference convolutions.py:241-250) as a native launch."""
    sd = {"conv.weight": conv.weight.detach(), "conv.bias": conv.bias.detach() if conv.bias is not None else None}
    return conv_block_ncdhw(x, sd, conv.kernel_size, conv.stride, False, False, "none")

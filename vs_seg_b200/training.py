"""Train-mode forward and backward of UNet2d5_spvPA on the native kernels (no torch/cuDNN compute).

Mirrors the module tree of the reference (/root/reference/params/networks/nets/unet2d5_spvPA.py:56-202)
as a tape of native launches:

  Convolution (convolutions.py:148-156, Conv -> BatchNorm3d -> Dropout -> PReLU)
      conv (tcgen05 kernel, bias in the epilogue) -> bn_stats -> bn_finalize (batch mean / biased var,
      running-stat update) -> bn_act_fwd (normalise + dropout + PReLU [+ ResidualUnit shortcut])
      backward: bn_act_bwd_reduce -> bn_act_bwd_apply -> conv3d_wgrad (+ bias grad) -> data gradient =
      the adjoint convolution on the SAME tcgen05 kernel (Conv3d <-> ConvTranspose3d with the same weights)
  AttentionBlock1/2 (attentionblock.py:10-47): conv+ReLU, conv+Sigmoid, gate x*(1+att) and their backward
  SkipConnection: channel views of one buffer (forward) / of one gradient buffer (backward)

torch only owns memory, provides the autograd boundary (`_UNetTrainFn`) and sums two 1-channel
attention-map gradients; parameters stay the module's own tensors, so torch.optim.Adam applies.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import lib as _lib
from .engine import _round_up, pack_conv_weight, pack_conv_weight_tc
from .tensors import Act8Buffer, f32view

_BN_EPS, _BN_MOM = 1e-5, 0.1


def _lib_tc_wgrad():
    import os
    return os.environ.get("VSSEG_TC_WGRAD", "1") != "0"


_SEED_BASE: dict = {}


def dropout_seed_base(device) -> torch.Tensor:
    """Per-device int64 scalar the BN/activation kernels add to their per-layer dropout salt.  It is rewritten before every
    step (eagerly by UNetTrainStep, or by GraphedTrainStep before a replay), so captured launches draw fresh masks."""
    key = str(torch.device(device))
    t = _SEED_BASE.get(key)
    if t is None:
        t = _SEED_BASE[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return t


class _GradBuf:
    """Gradient of an activation tensor: first writer stores, later writers accumulate."""

    def __init__(self, like: Act8Buffer):
        self.buf = Act8Buffer(like.B, like.C, like.X, like.Y, like.Z, like.device)
        self.filled = False


class UNetTrainStep:
    def __init__(self, model, x: torch.Tensor):
        self.lib = _lib.load()
        self.m = model
        self.dev = x.device
        self.B = x.shape[0]
        self.stream = torch.cuda.current_stream(self.dev).cuda_stream
        self.tape = []
        self._consts = {}
        # PReLU slopes are read by the kernels through device pointers (act.weight): the step never syncs the host
        self.keep = []
        self.pgrads = {}          # parameter name -> fp32 grad tensor (torch layout)
        self.p = dict(model.named_parameters())
        self.bufs = dict(model.named_buffers())
        self.drop_p = float(model.dropout or 0.0)
        # dropout masks = hash(per-layer salt + device-side base, element): the base is redrawn every step; while a CUDA
        # graph is being captured the launch only records the pointer (GraphedTrainStep stores a new base per replay)
        self.seed = 0
        self.seed_base = dropout_seed_base(self.dev)
        if not torch.cuda.is_current_stream_capturing():
            self.seed_base.fill_(int(torch.randint(0, 2 ** 62, (1,)).item()))
        self.layer_id = 0

    # ---- small helpers --------------------------------------------------------------------------------
    def _chk(self, code, what):
        if code:
            _lib.check(code, what)
        _lib.count_launch()

    def _buf(self, C_, dims):
        b = Act8Buffer(self.B, C_, dims[0], dims[1], dims[2], self.dev)
        self.keep.append(b)
        return b

    def _geom(self, k, s=(1, 1, 1), transposed=False):
        return _lib.ConvGeom(k[0], k[1], k[2], s[0], s[1], s[2], 1 if transposed else 0)

    def _addgrad(self, name, g):
        g = g.reshape(self.p[name].shape).to(self.p[name].dtype)
        self.pgrads[name] = g if name not in self.pgrads else self.pgrads[name] + g

    def _ident_ep(self, bias, cpad):
        scale = self._const(cpad, 1.0)
        shift = torch.nn.functional.pad(bias.detach().float(), (0, cpad - bias.numel())) if cpad != bias.numel() \
            else bias.detach().float()
        self.keep += [shift]
        return scale, shift

    # ---- convolution launches -------------------------------------------------------------------------
    def _const(self, n, value):
        """Cached all-ones / all-zeros epilogue vectors (one tiny launch per size instead of one per conv)."""
        key = (n, value)
        t = self._consts.get(key)
        if t is None:
            t = self._consts[key] = torch.full((n,), float(value), device=self.dev)
        return t

    def _pack_tc(self, w, transposed, adjoint, cin_pad, cout_pad, ns):
        """Weight image of vsseg_conv3d_tc built by ONE native launch (vsseg_pack_conv_weight_tc) instead of ~12
        torch launches.  w: the parameter in its own layout - Conv3d [Cout,Cin,k] (transposed=False), ConvTranspose3d
        [Cin,Cout,k] (transposed=True: sub-pixel phase image); adjoint=True: w is a stride-1 Conv3d weight and the
        image is that of its data-gradient conv (taps flipped, channel roles swapped)."""
        w = w.detach()
        if os.environ.get("VSSEG_NATIVE_PACK", "1") == "0":   # A/B switch: the torch-op packer this launch replaces
            if adjoint:
                w = w.flip(2, 3, 4).transpose(0, 1).contiguous()
            return pack_conv_weight_tc(_pad_cout(_pad_cin(w, transposed, cin_pad), transposed, cout_pad), transposed, ns)
        if not w.is_contiguous():
            w = w.contiguous()
        kx, ky, kz = w.shape[2:]
        conv_t, phases, flip = (1, 0, 1) if adjoint else ((1, 1, 0) if transposed else (0, 0, 0))
        nj = 2 if phases else kx
        n = 2 * ns * (2 if phases else 1) * (cin_pad // 16) * nj * kz * 2 * ky * (cout_pad // ns) * 8
        out = torch.empty(n, dtype=torch.bfloat16, device=self.dev)
        self._chk(self.lib.vsseg_pack_conv_weight_tc(w.data_ptr(), w.shape[0], w.shape[1], kx, ky, kz, conv_t, phases, flip,
                                                     cin_pad, cout_pad, ns, out.data_ptr(), self.stream), "pack_conv_weight_tc")
        return out

    def _conv(self, src, dst, geom, w_conv_layout, transposed, scale, shift, slope=1.0, act=0, res=None, adjoint=False):
        """dst = act(conv(src)*scale + shift) [+ res]; w in torch layout (Conv3d, or ConvTranspose3d if transposed;
        adjoint: the stride-1 Conv3d weight whose data-gradient conv this is)."""
        cpad = _round_up(dst.C, 16)
        ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), slope, act)
        res_p = C.byref(res) if res is not None else None
        ns = 0
        if src.C % 16 == 0:
            ns = self.lib.vsseg_conv3d_tc_suggest_split(C.byref(src), C.byref(dst), C.byref(geom), None)
        if ns > 0:
            w = self._pack_tc(w_conv_layout, transposed, adjoint, src.C, cpad, ns)
            self.keep.append(w)
            self._chk(self.lib.vsseg_conv3d_tc(C.byref(src), C.byref(dst), C.byref(geom), w.data_ptr(), ns, C.byref(ep),
                                               res_p, None, None, None, None, None, None, self.stream), "conv3d_tc")
        else:
            if adjoint:
                w_conv_layout = w_conv_layout.detach().flip(2, 3, 4).transpose(0, 1).contiguous()
            w_conv_layout = _pad_cin(w_conv_layout, transposed, src.C)   # zero weights for zero-padded input channels
            w = pack_conv_weight(w_conv_layout, transposed, cpad)
            self.keep.append(w)
            self._chk(self.lib.vsseg_conv3d_act8(C.byref(src), C.byref(dst), C.byref(geom), w.data_ptr(), cpad, C.byref(ep),
                                                 res_p, None, None, None, self.stream), "conv3d_act8")

    def _dgrad(self, dc, w, k, stride, transposed, gx: _GradBuf, c0=0, Cx=None):
        """Data gradient of a (transposed) conv = its adjoint convolution with the same weights."""
        dx = gx.buf.view(c0, Cx)
        cpad = _round_up(dx.C, 16)
        scale, shift = self._const(cpad, 1.0), self._const(cpad, 0.0)
        acc = gx.filled
        if not transposed and tuple(stride) == (1, 1, 1):
            # adjoint of a stride-1 Conv3d: Conv3d with the taps flipped and the channel roles swapped
            self._conv(dc, dx, self._geom(k), w, False, scale, shift, res=dx if acc else None, adjoint=True)
        elif not transposed:
            # strided Conv3d: adjoint = ConvTranspose3d whose weight tensor [in=Cout, out=Cin, k] is W itself
            self._conv(dc, dx, self._geom(k, stride, True), w.detach(), True, scale, shift, res=dx if acc else None)
        else:
            # ConvTranspose3d [Cin, Cout, k]: adjoint = strided Conv3d with weight [out=Cin, in=Cout, k] = W itself
            self._conv(dc, dx, self._geom(k, stride, False), w.detach(), False, scale, shift, res=dx if acc else None)
        gx.filled = True

    def _wgrad(self, x, dc, k, stride, transposed, wname, bname):
        w = self.p[wname]
        cout = w.shape[1] if transposed else w.shape[0]
        cin = w.shape[0] if transposed else w.shape[1]
        cpad = _round_up(dc.C, 16)
        taps = k[0] * k[1] * k[2]
        dw = torch.zeros((taps, x.C, cpad), device=self.dev)
        db = torch.zeros(cpad, device=self.dev)
        g = self._geom(k, stride, transposed)
        if _lib_tc_wgrad() and self.lib.vsseg_conv3d_wgrad_tc_supported(C.byref(x), C.byref(dc), C.byref(g)):
            # tensor-core weight gradient; the bias gradient is the plain channel sum of dc
            self._chk(self.lib.vsseg_conv3d_wgrad_tc(C.byref(x), C.byref(dc), C.byref(g), dw.data_ptr(), cpad, self.stream),
                      "conv3d_wgrad_tc")
            sums = torch.zeros(2 * dc.C, dtype=torch.float64, device=self.dev)
            self._chk(self.lib.vsseg_bn_stats(C.byref(dc), sums.data_ptr(), self.stream), "bn_stats")
            db[:dc.C] = sums[:dc.C].float()
        else:
            self._chk(self.lib.vsseg_conv3d_wgrad(C.byref(x), C.byref(dc), C.byref(g), dw.data_ptr(), cpad, db.data_ptr(),
                                                  self.stream), "conv3d_wgrad")
        dw = dw[:, :cin, :cout].reshape(k[0], k[1], k[2], cin, cout)
        self._addgrad(wname, dw.permute(3, 4, 0, 1, 2) if transposed else dw.permute(4, 3, 0, 1, 2))
        self._addgrad(bname, db[:cout])

    # ---- blocks ---------------------------------------------------------------------------------------
    def convolution(self, prefix, src, src_grad, dst, k, stride=(1, 1, 1), transposed=False, residual=None, cin1=None,
                    c0=0, Cx=None):
        """Conv -> BN(batch stats) -> Dropout -> PReLU [+ residual] into `dst` (an Act8 view).  src_grad: _GradBuf of the
        input (None: input needs no gradient); (c0, Cx) channel range of `src` inside src_grad.
        cin1: F32View of a 1-channel fp32 source (first conv) instead of `src`.  Returns a _GradBuf-less closure on the tape."""
        lib, s = self.lib, self.stream
        w, bias = self.p[prefix + "conv.weight"], self.p[prefix + "conv.bias"]
        gamma, beta = self.p[prefix + "norm.weight"], self.p[prefix + "norm.bias"]
        slope_t = self.p[prefix + "act.weight"]
        if slope_t.dtype != torch.float32 or slope_t.numel() != 1:
            raise NotImplementedError("native training covers PReLU with one shared fp32 slope (the reference's Act.PRELU)")
        slope = slope_t.data_ptr()
        Cc = dst.C
        cbuf = Act8Buffer(dst.B, Cc, dst.X, dst.Y, dst.Z, self.dev)
        self.keep.append(cbuf)
        c = cbuf.view()
        cpad = _round_up(Cc, 16)
        scale, shift = self._ident_ep(bias, cpad)
        geom = self._geom(k, stride, transposed)
        if cin1 is not None:
            wp = pack_conv_weight(w.detach(), False)[:, 0, :].contiguous()
            ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), 1.0, 0)
            self.keep += [wp]
            self._chk(lib.vsseg_conv3d_cin1(C.byref(cin1), C.byref(c), C.byref(geom), wp.data_ptr(), C.byref(ep), s), "cin1")
        else:
            self._conv(src, c, geom, w.detach(), transposed, scale, shift)
        count = dst.B * dst.X * dst.Y * dst.Z
        sums = torch.zeros(2 * Cc, dtype=torch.float64, device=self.dev)
        stats = torch.empty(4 * Cc, device=self.dev)
        rm, rv = self.bufs[prefix + "norm.running_mean"], self.bufs[prefix + "norm.running_var"]
        self._chk(lib.vsseg_bn_stats(C.byref(c), sums.data_ptr(), s), "bn_stats")
        self._chk(lib.vsseg_bn_finalize(sums.data_ptr(), Cc, count, gamma.data_ptr(), beta.data_ptr(), _BN_EPS, _BN_MOM,
                                        rm.data_ptr(), rv.data_ptr(), stats.data_ptr(), s), "bn_finalize")
        self.bufs[prefix + "norm.num_batches_tracked"].add_(1)
        self.layer_id += 1
        seed = (self.seed + 0x1000003 * self.layer_id) & (2 ** 63 - 1)
        res_p = C.byref(residual) if residual is not None else None
        sb = self.seed_base.data_ptr()
        self._chk(lib.vsseg_bn_act_fwd(C.byref(c), C.byref(dst), stats.data_ptr(), slope, self.drop_p, seed, sb, res_p, s), "bn_act_fwd")
        self.keep += [sums, stats, c, dst, src, residual]

        def backward(dy):
            sums2 = torch.zeros(2 * Cc + 1, dtype=torch.float64, device=self.dev)
            self._chk(lib.vsseg_bn_act_bwd_reduce(C.byref(c), C.byref(dy), stats.data_ptr(), slope, self.drop_p, seed, sb,
                                                  sums2.data_ptr(), s), "bn_act_bwd_reduce")
            self._addgrad(prefix + "norm.bias", sums2[:Cc].float())
            self._addgrad(prefix + "norm.weight", sums2[Cc:2 * Cc].float())
            self._addgrad(prefix + "act.weight", sums2[2 * Cc:].float())
            dcb = Act8Buffer(dst.B, Cc, dst.X, dst.Y, dst.Z, self.dev)
            dc = dcb.view()
            self._chk(lib.vsseg_bn_act_bwd_apply(C.byref(c), C.byref(dy), stats.data_ptr(), sums2.data_ptr(), slope, self.drop_p,
                                                 seed, sb, C.byref(dc), s), "bn_act_bwd_apply")
            if cin1 is not None:
                taps = k[0] * k[1] * k[2]
                dw = torch.zeros((taps, Cc), device=self.dev)
                db = torch.zeros(Cc, device=self.dev)
                self._chk(lib.vsseg_conv3d_cin1_wgrad(C.byref(cin1), C.byref(dc), C.byref(geom), dw.data_ptr(), db.data_ptr(), s),
                          "cin1_wgrad")
                self._addgrad(prefix + "conv.weight", dw.reshape(k[0], k[1], k[2], 1, Cc).permute(4, 3, 0, 1, 2))
                self._addgrad(prefix + "conv.bias", db)
                return
            self._wgrad(src, dc, k, stride, transposed, prefix + "conv.weight", prefix + "conv.bias")
            if src_grad is not None:
                self._dgrad(dc, w, k, stride, transposed, src_grad, c0, Cx)

        return backward

    def shortcut(self, prefix, src, src_grad, dst, cin1=None):
        """1x1x1 ResidualUnit shortcut conv (convolutions.py:241-250) into `dst`; returns its backward."""
        lib, s = self.lib, self.stream
        w, bias = self.p[prefix + "weight"], self.p[prefix + "bias"]
        cpad = _round_up(dst.C, 16)
        scale, shift = self._ident_ep(bias, cpad)
        geom = self._geom((1, 1, 1))
        if cin1 is not None:
            wp = pack_conv_weight(w.detach(), False)[:, 0, :].contiguous()
            ep = _lib.Epilogue(scale.data_ptr(), shift.data_ptr(), 1.0, 0)
            self.keep += [wp]
            self._chk(lib.vsseg_conv3d_cin1(C.byref(cin1), C.byref(dst), C.byref(geom), wp.data_ptr(), C.byref(ep), s), "cin1")
        else:
            self._conv(src, dst, geom, w.detach(), False, scale, shift)

        def backward(dy):
            if cin1 is not None:
                dw = torch.zeros((1, dst.C), device=self.dev)
                db = torch.zeros(dst.C, device=self.dev)
                self._chk(lib.vsseg_conv3d_cin1_wgrad(C.byref(cin1), C.byref(dy), C.byref(geom), dw.data_ptr(), db.data_ptr(), s),
                          "cin1_wgrad")
                self._addgrad(prefix + "weight", dw.reshape(1, 1, 1, 1, dst.C).permute(4, 3, 0, 1, 2))
                self._addgrad(prefix + "bias", db)
                return
            self._wgrad(src, dy, (1, 1, 1), (1, 1, 1), False, prefix + "weight", prefix + "bias")
            if src_grad is not None:
                self._dgrad(dy, w, (1, 1, 1), (1, 1, 1), False, src_grad)

        return backward

    def residual_unit(self, prefix, src, src_grad, dst_buf, dst_c0, cout, k, subunits, cin1=None):
        """ResidualUnit (convolutions.py:209-255): out = conv path + shortcut, written to dst_buf[dst_c0 : dst_c0+cout].
        Returns backward(dout_view)."""
        dims = (dst_buf.X, dst_buf.Y, dst_buf.Z)
        sbuf = self._buf(cout, dims)
        bw_sc = self.shortcut(prefix + "residual.", src, src_grad, sbuf.view(), cin1=cin1)
        dst = dst_buf.view(dst_c0, cout)
        if subunits == 2:
            hbuf = self._buf(cout, dims)
            gh = _GradBuf(hbuf)
            bw0 = self.convolution(prefix + "conv.unit0.", src, src_grad, hbuf.view(), k, cin1=cin1)
            bw1 = self.convolution(prefix + "conv.unit1.", hbuf.view(), gh, dst, k, residual=sbuf.view())

            def backward(dout):
                bw1(dout)
                bw0(gh.buf.view())
                bw_sc(dout)
        else:
            bw0 = self.convolution(prefix + "conv.unit0.", src, src_grad, dst, k, residual=sbuf.view())

            def backward(dout):
                bw0(dout)
                bw_sc(dout)
        return backward

    def attention(self, prefix, xbuf: Act8Buffer, gx: _GradBuf, k):
        """AttentionBlock1 + AttentionBlock2 on all channels of xbuf; returns (gated buffer, its _GradBuf, att tensor,
        backward(datt_from_loss))."""
        lib, s = self.lib, self.stream
        Cc = xbuf.C
        dims = (xbuf.X, xbuf.Y, xbuf.Z)
        x = xbuf.view()
        hC = _round_up(Cc // 2, 16)
        hbuf = self._buf(hC, dims)
        h = hbuf.view()
        w1, b1 = self.p[prefix + "0.conv1.conv.weight"], self.p[prefix + "0.conv1.conv.bias"]
        w2, b2 = self.p[prefix + "0.conv2.conv.weight"], self.p[prefix + "0.conv2.conv.bias"]
        sc1, sh1 = self._ident_ep(b1, _round_up(hC, 16))
        self._conv(x, h, self._geom(k), w1.detach(), False, sc1, sh1, slope=0.0)        # conv + ReLU
        att = torch.empty((self.B, 1) + dims, device=self.dev)
        av = f32view(att)
        w2p = w2.detach().float()
        if w2p.shape[1] != hC:
            w2p = torch.nn.functional.pad(w2p, (0, 0, 0, 0, 0, 0, 0, hC - w2p.shape[1]))
        w2g = pack_conv_weight(w2p, False)                                               # [taps][hC][1]
        b2f = b2.detach().float().contiguous()
        g2 = self._geom(k)
        self._chk(lib.vsseg_conv3d_smallcout(C.byref(h), C.byref(av), C.byref(g2), w2g.data_ptr(), b2f.data_ptr(), 1, 0.0, None, s),
                  "smallcout")
        gbuf = self._buf(Cc, dims)
        g = gbuf.view()
        self._chk(lib.vsseg_att_gate(C.byref(x), C.byref(av), C.byref(g), s), "att_gate")
        gg = _GradBuf(gbuf)
        self.keep += [w2g, b2f, att, av, x, h, g]

        def backward(datt_loss):
            # gate: dx = dg*(1+att), datt = sum_c dg*x
            datt = torch.empty_like(att)
            dav = f32view(datt)
            dx = gx.buf.view()
            self._chk(lib.vsseg_att_gate_bwd(C.byref(x), C.byref(av), C.byref(gg.buf.view()), C.byref(dx), C.byref(dav),
                                             1 if gx.filled else 0, s), "att_gate_bwd")
            gx.filled = True
            if datt_loss is not None:
                datt = datt + datt_loss.reshape(datt.shape).float()
                dav = f32view(datt)
            # conv2 + sigmoid backward -> dh, dW2, db2
            gh = _GradBuf(hbuf)
            taps = k[0] * k[1] * k[2]
            dw2 = torch.zeros((taps, hC, 1), device=self.dev)
            db2 = torch.zeros(1, device=self.dev)
            dh = gh.buf.view()
            self._chk(lib.vsseg_conv3d_smallcout_bwd(C.byref(h), C.byref(dav), C.byref(av), C.byref(g2), w2g.data_ptr(), 1,
                                                     C.byref(dh), dw2.data_ptr(), db2.data_ptr(), s), "smallcout_bwd")
            cin2 = w2.shape[1]
            self._addgrad(prefix + "0.conv2.conv.weight",
                          dw2[:, :cin2, :].reshape(k[0], k[1], k[2], cin2, 1).permute(4, 3, 0, 1, 2))
            self._addgrad(prefix + "0.conv2.conv.bias", db2)
            # conv1 + ReLU backward
            dcb = Act8Buffer(self.B, hC, *dims, self.dev)
            dc = dcb.view()
            self._chk(lib.vsseg_act_bwd(C.byref(h), C.byref(dh), 0.0, C.byref(dc), s), "act_bwd")
            self._wgrad(x, dc, k, (1, 1, 1), False, prefix + "0.conv1.conv.weight", prefix + "0.conv1.conv.bias")
            self._dgrad(dc, w1, k, (1, 1, 1), False, gx)
            self.keep += [datt, dcb, gh]

        return gbuf, gg, att, backward

    # ---- whole network ----------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor):
        m = self.m
        lib, s = self.lib, self.stream
        ch = tuple(m.channels)
        nlev = len(ch)
        strides = [tuple(v) for v in m.strides]
        ks = [tuple(v) for v in m.kernel_sizes]
        sks = [tuple(v) for v in m.sample_kernel_sizes]
        att_on = bool(m.attention_module)
        dims = [tuple(x.shape[2:])]
        for st in strides:
            dims.append(tuple(d // q for d, q in zip(dims[-1], st)))
        xin = x.detach().float().contiguous()
        self.keep.append(xin)
        xv = f32view(xin)
        cat = [self._buf(2 * ch[l], dims[l]) for l in range(nlev - 1)]
        gcat = [_GradBuf(cb) for cb in cat]
        dn = [self._buf(ch[l], dims[l + 1]) for l in range(nlev - 1)]
        gdn = [_GradBuf(b) for b in dn]
        prefixes, p = [], "model."
        for l in range(nlev - 1):
            prefixes.append(p)
            p = p + "1.submodule.1."
        tape = self.tape
        atts = []          # (att tensor, backward) in hook order: coarsest first
        # ---- encoder
        for l in range(nlev - 1):
            pr = prefixes[l]
            if l == 0:
                bw = self.residual_unit(pr + "0.", None, None, cat[0], 0, ch[0], ks[0], 2, cin1=xv)
            else:
                bw = self.residual_unit(pr + "0.", dn[l - 1].view(), gdn[l - 1], cat[l], 0, ch[l], ks[l], 2)
            tape.append((bw, lambda l=l: gcat[l].buf.view(0, ch[l])))
            bwd = self.convolution(pr + "1.submodule.0.", cat[l].view(0, ch[l]), gcat[l], dn[l].view(), sks[l], strides[l],
                                   c0=0, Cx=ch[l])
            tape.append((bwd, lambda l=l: gdn[l].buf.view()))
        # ---- bottom
        pb = prefixes[-1] + "1.submodule.1."
        kb = ks[-1]
        bot = self._buf(ch[-1], dims[-1])
        gbot = _GradBuf(bot)
        if att_on:
            gbuf, gg, att, bwa = self.attention(pb + "0.", dn[-1], gdn[-1], kb)
            atts.append((att, bwa))
            bwr = self.residual_unit(pb + "1.", gbuf.view(), gg, bot, 0, ch[-1], kb, 2)
            tape.append((bwa, None))
            tape.append((bwr, lambda: gbot.buf.view()))
        else:
            bwr = self.residual_unit(pb, dn[-1].view(), gdn[-1], bot, 0, ch[-1], kb, 2)
            tape.append((bwr, lambda: gbot.buf.view()))
        # ---- decoder
        sub, gsub = bot, gbot
        logits = None
        for l in range(nlev - 2, -1, -1):
            pr = prefixes[l]
            bwu = self.convolution(pr + "1.submodule.2.", sub.view(), gsub, cat[l].view(ch[l], ch[l]), sks[l], strides[l],
                                   transposed=True)
            tape.append((bwu, lambda l=l: gcat[l].buf.view(ch[l], ch[l])))
            pu = pr + "2."
            src_buf, src_g = cat[l], gcat[l]
            if att_on:
                gbuf, gg, att, bwa = self.attention(pu + "0.", cat[l], gcat[l], ks[l])
                atts.append((att, bwa))
                tape.append((bwa, None))
                src_buf, src_g = gbuf, gg
                pu = pu + "1."
            if l > 0:
                out = self._buf(ch[l], dims[l])   # the decoder ResidualUnit at level l maps 2*ch[l] -> ch[l]
                gout = _GradBuf(out)
                bwr = self.residual_unit(pu, src_buf.view(), src_g, out, 0, ch[l], ks[l], 1)
                tape.append((bwr, lambda gout=gout: gout.buf.view()))
                sub, gsub = out, gout
            else:
                # top unit: conv_only + 1x1x1 shortcut, both linear: one small-Cout conv with the shortcut folded
                # into the centre tap; the parameter gradients are un-folded in backward
                k0 = ks[0]
                nout = m.out_channels
                wc, bc = self.p[pu + "conv.unit0.conv.weight"], self.p[pu + "conv.unit0.conv.bias"]
                wr, br = self.p[pu + "residual.weight"], self.p[pu + "residual.bias"]
                wf = wc.detach().float().clone()
                wf[:, :, k0[0] // 2, k0[1] // 2, k0[2] // 2] += wr.detach().reshape(nout, -1).float()
                wg = pack_conv_weight(wf, False)                       # [taps][Cin][nout]
                bf = (bc.detach() + br.detach()).float().contiguous()
                logits = torch.empty((self.B, nout) + dims[0], device=self.dev)
                lv = f32view(logits)
                g0 = self._geom(k0)
                sv = src_buf.view()
                self._chk(lib.vsseg_conv3d_smallcout(C.byref(sv), C.byref(lv), C.byref(g0), wg.data_ptr(), bf.data_ptr(), 0, 1.0,
                                                     None, s), "logits")
                self.keep += [wg, bf, sv, lv]

                def bw_top(dlogits, sv=sv, src_g=src_g, wg=wg, g0=g0, k0=k0, pu=pu, nout=nout, wc=wc):
                    dl = dlogits.float().contiguous()
                    dlv = f32view(dl)
                    taps = k0[0] * k0[1] * k0[2]
                    dw = torch.zeros((taps, sv.C, nout), device=self.dev)
                    db = torch.zeros(nout, device=self.dev)
                    dx = src_g.buf.view()
                    if src_g.filled:
                        raise RuntimeError("top unit must be the first writer of its input gradient")
                    self._chk(lib.vsseg_conv3d_smallcout_bwd(C.byref(sv), C.byref(dlv), None, C.byref(g0), wg.data_ptr(), 0,
                                                             C.byref(dx), dw.data_ptr(), db.data_ptr(), s), "logits_bwd")
                    src_g.filled = True
                    cin = wc.shape[1]
                    dwt = dw[:, :cin, :].reshape(k0[0], k0[1], k0[2], cin, nout).permute(4, 3, 0, 1, 2)
                    self._addgrad(pu + "conv.unit0.conv.weight", dwt)
                    self._addgrad(pu + "residual.weight", dwt[:, :, k0[0] // 2, k0[1] // 2, k0[2] // 2].reshape(nout, cin, 1, 1, 1))
                    self._addgrad(pu + "conv.unit0.conv.bias", db)
                    self._addgrad(pu + "residual.bias", db)
                    self.keep += [dl, dw, db]

                self.bw_top = bw_top
        self.atts = atts
        return logits, [a for a, _ in atts]

    def backward(self, dlogits, datts):
        """Runs the tape in reverse; returns {parameter name: gradient}."""
        att_grads = {}
        for (att, bwa), g in zip(self.atts, datts):
            att_grads[id(bwa)] = g
        self.bw_top(dlogits)
        for bw, grad_view in reversed(self.tape):
            if grad_view is None:
                bw(att_grads.get(id(bw)))
            else:
                bw(grad_view())
        return self.pgrads


def _pad_cin(w, transposed, c):
    d = 0 if transposed else 1
    if w.shape[d] == c:
        return w
    pad = [0, 0] * (w.dim() - 1 - d) + [0, c - w.shape[d]]
    return torch.nn.functional.pad(w, pad)


def _pad_cout(w, transposed, c):
    d = 1 if transposed else 0
    if w.shape[d] == c:
        return w
    pad = [0, 0] * (w.dim() - 1 - d) + [0, c - w.shape[d]]
    return torch.nn.functional.pad(w, pad)


class _UNetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, names, *params):
        step = UNetTrainStep(model, x)
        logits, atts = step.forward(x)
        ctx.step, ctx.names = step, names
        return (logits, *atts)

    @staticmethod
    def backward(ctx, dlogits, *datts):
        step = ctx.step
        if dlogits is None:
            raise RuntimeError("the logits received no gradient")
        grads = step.backward(dlogits, list(datts))
        out = [grads.get(n) for n in ctx.names]
        ctx.step = None
        return (None, None, None, *out)


def unet_train_forward(model, x):
    """Train-mode forward of UNet2d5_spvPA on CUDA through the native kernels; differentiable w.r.t. every
    parameter of `model` (BatchNorm running statistics are updated in place)."""
    if not x.is_cuda:
        raise _lib.NativeLibraryError("unet_train_forward needs a CUDA tensor (no CPU fallback)")
    if not model._plan_supported():
        raise NotImplementedError("the native training path covers the reference configuration "
                                  "(3-D, num_res_units=2, BatchNorm, PReLU, 1 input channel)")
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    names = tuple(n for n, _ in named)
    outs = _UNetTrainFn.apply(model, x, names, *[p for _, p in named])
    return outs[0], list(outs[1:])


# ---- standalone blocks in train mode (reference convolutions.py:148-156, :209-255 under autograd) ------------------
class _BlockHost:
    """What UNetTrainStep needs from its `model` when the model is a single Convolution / ResidualUnit."""

    def __init__(self, module, drop_p):
        self._m, self.dropout = module, drop_p

    def named_parameters(self):
        return self._m.named_parameters()

    def named_buffers(self):
        return self._m.named_buffers()


def _block_drop_p(module):
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            return float(m.p)
    return 0.0


class _BlockTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, kind, names, *params):
        B, cin = x.shape[:2]
        step = UNetTrainStep(_BlockHost(module, _block_drop_p(module)), x)
        src = Act8Buffer(B, cin, *x.shape[2:], x.device).from_ncdhw(x.detach())
        gsrc = _GradBuf(src) if ctx.needs_input_grad[1] else None
        if kind == "conv":
            c = module.conv
            cout = c.out_channels
            odims = [d * s for d, s in zip(x.shape[2:], c.stride)] if module.is_transposed else \
                [(d + s - 1) // s for d, s in zip(x.shape[2:], c.stride)]
            dst = Act8Buffer(B, cout, *odims, x.device)
            bw = step.convolution("", src.view(), gsrc, dst.view(), tuple(c.kernel_size), tuple(c.stride), module.is_transposed)
        else:
            units = list(module.conv.children())
            c = units[0].conv
            cout = module.out_channels
            dst = Act8Buffer(B, cout, *x.shape[2:], x.device)
            bw = step.residual_unit("", src.view(), gsrc, dst, 0, cout, tuple(c.kernel_size), len(units))
        ctx.step, ctx.bw, ctx.gsrc, ctx.names, ctx.oshape = step, bw, gsrc, names, (B, cout, dst.X, dst.Y, dst.Z)
        ctx.keep = (src, dst)
        return dst.to_ncdhw()

    @staticmethod
    def backward(ctx, gy):
        B, cout, X, Y, Z = ctx.oshape
        dout = Act8Buffer(B, cout, X, Y, Z, gy.device).from_ncdhw(gy.contiguous())
        ctx.bw(dout.view())
        grads = ctx.step.pgrads
        gx = ctx.gsrc.buf.to_ncdhw() if ctx.gsrc is not None else None
        out = [grads.get(n) for n in ctx.names]
        ctx.step = ctx.bw = None
        return (None, gx, None, None, *out)


def block_train_forward(module, x):
    """Train-mode forward of a standalone ``Convolution`` (Conv -> BatchNorm3d(batch statistics) -> Dropout -> PReLU)
    or ``ResidualUnit`` on CUDA through the native training kernels, differentiable w.r.t. the block's parameters and
    its input.  Covers the block types the network is made of: 3-D, BatchNorm + PReLU, channels % 8 == 0, stride-1
    ResidualUnits with a 1x1x1 shortcut; anything else raises (there is no eager CUDA fallback)."""
    from params.networks.blocks.convolutions import Convolution, ResidualUnit
    if not x.is_cuda:
        raise _lib.NativeLibraryError("block_train_forward needs a CUDA tensor (no CPU fallback)")

    def conv_ok(m):
        return (isinstance(m, Convolution) and m._native_supported() and m._norm_name == "BATCH" and m._act_name == "PRELU"
                and m.conv.in_channels % 8 == 0 and m.conv.out_channels % 8 == 0 and m.conv.bias is not None)

    if isinstance(module, Convolution):
        ok, kind = conv_ok(module), "conv"
    elif isinstance(module, ResidualUnit):
        units = list(module.conv.children())
        ok = (len(units) in (1, 2) and all(conv_ok(u) and tuple(u.conv.stride) == (1, 1, 1) and not u.is_transposed for u in units)
              and isinstance(module.residual, torch.nn.Conv3d) and tuple(module.residual.kernel_size) == (1, 1, 1))
        kind = "ru"
    else:
        ok, kind = False, ""
    if not ok:
        raise NotImplementedError("native train-mode blocks: Convolution / ResidualUnit with BatchNorm + PReLU, "
                                  "channels % 8 == 0 (ResidualUnit: stride 1, 1x1x1 shortcut)")
    named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
    return _BlockTrainFn.apply(module, x.float(), kind, tuple(n for n, _ in named), *[p for _, p in named])


# ---- the whole training step as one CUDA graph -----------------------------------------------------------------------
class GraphedTrainStep:
    """``loss = step(inputs, labels)`` = the reference's inner loop body (VSparams.py:457-462: zero_grad, forward, loss,
    backward, optimizer.step) captured ONCE per input shape in a CUDA graph and replayed.

    The eager step issues ~420 native launches plus the Python tape around them (~40 ms of host time at best, several
    times that when the host cores are shared), which made the step host-bound and the data-parallel step scale badly.
    A replay costs the host four calls: copy the batch into the graph's static input tensors, store a fresh dropout seed
    base and the Adam step count / learning rate in device scalars (the captured kernels read them there), launch.
    BatchNorm running statistics, the flat Adam moments and, under torchrun, the NCCL all-reduce of the flat gradient are
    all part of the graph.  Parameters must belong to a ``FusedAdam`` on CUDA.
    """

    def __init__(self, model, loss_function, optimizer, reducer=None, warmup=2):
        self.model, self.loss_function, self.optimizer, self.reducer = model, loss_function, optimizer, reducer
        self.warmup = int(warmup)
        self._graphs = {}
        if not hasattr(optimizer, "pre_replay"):
            raise TypeError("GraphedTrainStep needs a vs_seg_b200.optim.FusedAdam")

    def _eager(self, x, y):
        self.optimizer.zero_grad()
        loss = self.loss_function(self.model(x), y)
        loss.backward()
        if self.reducer is not None:
            self.reducer.reduce()
        self.optimizer.step()
        return loss.detach()

    def _snapshot(self):
        fl = [f for f in self.optimizer._flat if f is not None]
        return ([(f, f["p"].clone(), f["m"].clone(), f["v"].clone()) for f in fl],
                [(b, b.clone()) for b in self.model.buffers()], [g["step"] for g in self.optimizer.param_groups])

    def _restore(self, snap):
        flat, bufs, steps = snap
        with torch.no_grad():
            for f, p, m, v in flat:
                f["p"].copy_(p)
                f["m"].copy_(m)
                f["v"].copy_(v)
            for b, c in bufs:
                b.copy_(c)
        for g, st in zip(self.optimizer.param_groups, steps):
            g["step"] = st

    def _capture(self, inputs, labels):
        import gc
        dev = inputs.device
        sx, sy = inputs.detach().clone(), labels.detach().clone()
        # autograd caches one AccumulateGrad node per parameter for as long as a graph that uses it is alive, and runs it
        # on the stream it was created on: nodes left over from earlier eager steps (kept alive by model.att_maps or by a
        # loss the caller still holds) would pull the legacy stream into the capture and invalidate it.  Drop our own
        # reference, and run warm-up and capture on ONE side stream so the fresh nodes belong to the capture's stream.
        self.model.att_maps = []
        gc.collect()
        # warm-up passes (lazy initialisation, allocator pools, function attributes must not happen inside the capture)
        # are real steps on this batch, so the training state is put back afterwards
        snap = self._snapshot()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._eager(sx, sy)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snap)
        self.model.att_maps = []
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            loss = self._eager(sx, sy)
        self.model.att_maps = []   # the captured forward's maps are graph-private memory: do not hand them out
        return {"graph": graph, "x": sx, "y": sy, "loss": loss}

    def __call__(self, inputs, labels):
        if not inputs.is_cuda:
            raise _lib.NativeLibraryError("GraphedTrainStep runs on CUDA tensors only")
        key = (tuple(inputs.shape), inputs.dtype, tuple(labels.shape), labels.dtype, str(inputs.device))
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 2:   # every graph keeps the activations of a whole step alive
                self._graphs.clear()
            g = self._graphs[key] = self._capture(inputs, labels)
        g["x"].copy_(inputs, non_blocking=True)
        g["y"].copy_(labels, non_blocking=True)
        dropout_seed_base(inputs.device).fill_(int(torch.randint(0, 2 ** 62, (1,)).item()))
        self.optimizer.pre_replay()
        g["graph"].replay()
        self.optimizer.post_replay()
        _lib.count_launch()
        return g["loss"]

"""Oracle: functional fp32 restatement of UNet2d5_spvPA (test infrastructure only).

Works directly on a reference-format ``state_dict`` (the 256 keys of
``/root/reference/params/networks/nets/unet2d5_spvPA.py``), so it shares no
code with the product modules in ``params/``.

Reference semantics followed:
  * block order Conv -> BatchNorm -> Dropout -> PReLU
    (reference params/networks/blocks/convolutions.py:148-156)
  * ResidualUnit = conv path + 1x1x1 shortcut conv, no activation after the sum
    (convolutions.py:241-255)
  * transposed conv with output_padding = stride - 1 (convolutions.py:114-135)
  * attention gate conv(C->C/2,k)+ReLU, conv(C/2->1,k)+Sigmoid, x*(1+att)
    (params/networks/blocks/attentionblock.py:10-47)
  * recursive assembly, skip channels first in the concat
    (unet2d5_spvPA.py:56-89; MONAI SkipConnection = cat([x, sub(x)], 1))
  * att_maps returned coarsest first (hook order, unet2d5_spvPA.py:95-104)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# Hyper-parameters hard-coded by the reference (params/VSparams.py:343-374).
CHANNELS = (16, 32, 48, 64, 80, 96)
STRIDES = ((2, 2, 1), (2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2))
KERNEL_SIZES = ((3, 3, 1), (3, 3, 1), (3, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))
SAMPLE_KERNEL_SIZES = ((3, 3, 1), (3, 3, 1), (3, 3, 3), (3, 3, 3), (3, 3, 3))


def _pad(k):
    return tuple((int(v) - 1) // 2 for v in k)


def _convolution(sd, p, x, k, stride=(1, 1, 1), transposed=False, act="prelu", norm=True,
                 conv_only=False, training=False):
    """One reference ``Convolution`` block (convolutions.py:22-156), dropout = identity."""
    w, b = sd[p + "conv.weight"], sd[p + "conv.bias"]
    pad = _pad(k)
    if transposed:
        opad = tuple(s + 2 * q - (kk - 1) - 1 for s, q, kk in zip(stride, pad, k))
        y = F.conv_transpose3d(x, w, b, stride=stride, padding=pad, output_padding=opad)
    else:
        y = F.conv3d(x, w, b, stride=stride, padding=pad)
    if conv_only:
        return y
    if norm:
        y = F.batch_norm(y, sd[p + "norm.running_mean"], sd[p + "norm.running_var"],
                         sd[p + "norm.weight"], sd[p + "norm.bias"],
                         training=training, momentum=0.1, eps=1e-5)
    if act == "prelu":
        y = F.prelu(y, sd[p + "act.weight"])
    elif act == "relu":
        y = F.relu(y)
    elif act == "sigmoid":
        y = torch.sigmoid(y)
    return y


def _residual_unit(sd, p, x, k, subunits, last_conv_only=False, training=False):
    """Reference ``ResidualUnit.forward`` (convolutions.py:252-255)."""
    res = F.conv3d(x, sd[p + "residual.weight"], sd[p + "residual.bias"])
    cx = x
    for su in range(subunits):
        conv_only = last_conv_only and su == subunits - 1
        cx = _convolution(sd, f"{p}conv.unit{su}.", cx, k, conv_only=conv_only, training=training)
    return cx + res


def _att_gate(sd, p, x, k, att_maps):
    """AttentionBlock1 + AttentionBlock2 (attentionblock.py:32-47); p ends in '0.'."""
    h = _convolution(sd, p + "0.conv1.", x, k, act="relu", norm=False)
    att = _convolution(sd, p + "0.conv2.", h, k, act="sigmoid", norm=False)
    att_maps.append(att)
    return att.repeat(1, x.shape[1], 1, 1, 1) * x + x


def _level(sd, p, x, lvl, n_levels, attention, att_maps, training, kernel_sizes, strides,
           sample_kernel_sizes):
    """One recursion level of ``_create_block`` (unet2d5_spvPA.py:56-89)."""
    k, s, sk = kernel_sizes[lvl], strides[lvl], sample_kernel_sizes[lvl]
    e = _residual_unit(sd, p + "0.", x, k, subunits=2, training=training)
    d = _convolution(sd, p + "1.submodule.0.", e, sk, stride=s, training=training)
    if lvl + 2 < n_levels:
        sub = _level(sd, p + "1.submodule.1.", d, lvl + 1, n_levels, attention, att_maps, training,
                     kernel_sizes, strides, sample_kernel_sizes)
    else:  # bottom layer (unet2d5_spvPA.py:152-158)
        kb = kernel_sizes[lvl + 1]
        if attention:
            g = _att_gate(sd, p + "1.submodule.1.0.", d, kb, att_maps)
            sub = _residual_unit(sd, p + "1.submodule.1.1.", g, kb, subunits=2, training=training)
        else:
            sub = _residual_unit(sd, p + "1.submodule.1.", d, kb, subunits=2, training=training)
    u = _convolution(sd, p + "1.submodule.2.", sub, sk, stride=s, transposed=True, training=training)
    cat = torch.cat([e, u], dim=1)
    is_top = lvl == 0
    if attention:
        g = _att_gate(sd, p + "2.0.", cat, k, att_maps)
        return _residual_unit(sd, p + "2.1.", g, k, subunits=1, last_conv_only=is_top, training=training)
    return _residual_unit(sd, p + "2.", cat, k, subunits=1, last_conv_only=is_top, training=training)


def unet_forward(sd, x, attention=True, training=False, channels=CHANNELS, strides=STRIDES,
                 kernel_sizes=KERNEL_SIZES, sample_kernel_sizes=SAMPLE_KERNEL_SIZES):
    """``UNet2d5_spvPA.forward`` (unet2d5_spvPA.py:204-206): returns (logits, att_maps).

    ``training=True`` uses batch statistics in BatchNorm (running stats in ``sd`` are
    updated in place as torch does); dropout is always the identity here (p=0).
    """
    att_maps = []
    y = _level(sd, "model.", x, 0, len(channels), attention, att_maps, training,
               kernel_sizes, strides, sample_kernel_sizes)
    return y, att_maps


# --------------------------------------------------------------------------------------
# Deterministic weights shared by the oracle, the golden generator and the product tests.
# --------------------------------------------------------------------------------------
def _conv_entry(sd, g, p, cin, cout, k, transposed=False):
    shape = (cin, cout, *k) if transposed else (cout, cin, *k)
    fan_in = (cout if transposed else cin) * math.prod(k)
    bound = 1.0 / math.sqrt(fan_in)
    sd[p + "weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound * math.sqrt(3.0)
    sd[p + "bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound


def _block_entries(sd, g, p, cin, cout, k, transposed=False, conv_only=False):
    _conv_entry(sd, g, p + "conv.", cin, cout, k, transposed)
    if conv_only:
        return
    sd[p + "norm.weight"] = 0.75 + 0.5 * torch.rand(cout, generator=g)
    sd[p + "norm.bias"] = 0.2 * torch.randn(cout, generator=g)
    sd[p + "norm.running_mean"] = 0.1 * torch.randn(cout, generator=g)
    sd[p + "norm.running_var"] = 0.5 + torch.rand(cout, generator=g)
    sd[p + "norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd[p + "act.weight"] = 0.05 + 0.4 * torch.rand(1, generator=g)


def _ru_entries(sd, g, p, cin, cout, k, subunits, last_conv_only=False):
    c = cin
    for su in range(subunits):
        _block_entries(sd, g, f"{p}conv.unit{su}.", c, cout, k,
                       conv_only=last_conv_only and su == subunits - 1)
        c = cout
    _conv_entry(sd, g, p + "residual.", cin, cout, (1, 1, 1))


def _att_entries(sd, g, p, c, k):
    _conv_entry(sd, g, p + "0.conv1.conv.", c, c // 2, k)
    _conv_entry(sd, g, p + "0.conv2.conv.", c // 2, 1, k)


def seeded_state_dict(seed=0, attention=True, in_channels=1, out_channels=2, channels=CHANNELS,
                      strides=STRIDES, kernel_sizes=KERNEL_SIZES,
                      sample_kernel_sizes=SAMPLE_KERNEL_SIZES):
    """A reference-format state_dict with seeded, non-trivial values.

    Conv weights follow the scale of torch's default init; BatchNorm affine, running
    statistics and PReLU slopes are perturbed so that eval mode is not an identity
    normalisation (SURVEY.md §8c golden vector (i)).  The key set equals the reference
    module's ``state_dict()`` (checked by oracle/make_golden.py via load_state_dict).
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def level(p, lvl, inc, outc):
        c, k, sk = channels[lvl], kernel_sizes[lvl], sample_kernel_sizes[lvl]
        _ru_entries(sd, g, p + "0.", inc, c, k, 2)
        _block_entries(sd, g, p + "1.submodule.0.", c, c, sk)
        if lvl + 2 < len(channels):
            level(p + "1.submodule.1.", lvl + 1, c, channels[lvl + 1])
        else:
            kb = kernel_sizes[lvl + 1]
            if attention:
                _att_entries(sd, g, p + "1.submodule.1.0.", c, kb)
                _ru_entries(sd, g, p + "1.submodule.1.1.", c, channels[lvl + 1], kb, 2)
            else:
                _ru_entries(sd, g, p + "1.submodule.1.", c, channels[lvl + 1], kb, 2)
        _block_entries(sd, g, p + "1.submodule.2.", channels[lvl + 1], c, sk, transposed=True)
        if attention:
            _att_entries(sd, g, p + "2.0.", 2 * c, k)
            _ru_entries(sd, g, p + "2.1.", 2 * c, outc, k, 1, last_conv_only=lvl == 0)
        else:
            _ru_entries(sd, g, p + "2.", 2 * c, outc, k, 1, last_conv_only=lvl == 0)

    level("model.", 0, in_channels, out_channels)
    return sd

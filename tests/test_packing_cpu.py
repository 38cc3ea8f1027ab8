"""Host-side weight images and epilogue constants of the tensor-core conv (no GPU needed): the packers follow the layout
include/vsseg_b200.h documents for vsseg_conv3d_tc, and the folded scale/shift reproduce conv bias + eval-mode
BatchNorm3d + PReLU of the reference's Convolution block (convolutions.py:148-156)."""
import random

import torch

from vs_seg_b200.engine import (fold_epilogue, pack_conv_weight_tc, pack_conv_weight_tc_2p, pack_shortcut_weight_tc)


def _val(img, *idx):
    """hi + lo of one element of a packed image [..., plane, ...]: idx carries None at the plane position."""
    i = list(idx)
    p = i.index(None)
    i[p] = 0
    hi = img[tuple(i)].float()
    i[p] = 1
    return (hi + img[tuple(i)].float()).item()


def test_conv_weight_image_follows_the_header_layout():
    g = torch.Generator().manual_seed(3)
    cout, cin, k = 40, 32, (3, 3, 3)          # Cout padded to 48; two N slices of 24 are not allowed, three of 16 are
    w = torch.randn((cout, cin) + k, generator=g)
    for n_split in (1, 3):
        img = pack_conv_weight_tc(w, False, n_split)
        n_cta = 48 // n_split
        assert img.dtype == torch.bfloat16 and tuple(img.shape) == (n_split, cin // 16, 3, 2, 3, 2, 3, n_cta, 8)
        rnd = random.Random(n_split)
        for _ in range(300):
            sel, c, j, tz, kh, typ, n, e = (rnd.randrange(v) for v in (n_split, cin // 16, 3, 3, 2, 3, n_cta, 8))
            co, ci, ty = sel * n_cta + n, 16 * c + 8 * kh + e, 2 - typ      # ty' = ky - 1 - ty
            want = w[co, ci, j, ty, tz].item() if co < cout else 0.0
            got = _val(img, sel, c, j, None, tz, kh, typ, n, e)
            assert abs(got - want) <= 2.0 ** -15 * abs(want) + 1e-30, (sel, c, j, tz, kh, typ, n, e)


def test_transposed_conv_weight_image_is_split_by_output_parity():
    g = torch.Generator().manual_seed(4)
    cin, cout = 32, 16
    w = torch.randn((cin, cout, 3, 3, 3), generator=g)      # ConvTranspose3d layout [Cin, Cout, kx, ky, kz]
    img = pack_conv_weight_tc(w, True, 1)
    assert tuple(img.shape) == (2, cin // 16, 2, 2, 3, 2, 3, 16, 8)   # sel = output x parity, j = input x shift
    rnd = random.Random(7)
    for _ in range(300):
        px, c, j, tz, kh, ty, n, e = (rnd.randrange(v) for v in (2, cin // 16, 2, 3, 2, 3, 16, 8))
        ci = 16 * c + 8 * kh + e
        kx = {(0, 0): 1, (1, 0): 2, (1, 1): 0}.get((px, j))       # even outputs: centre tap only; odd: taps 2 and 0
        want = w[ci, n, kx, ty, tz].item() if kx is not None else 0.0   # ty' = ty for the transposed conv
        got = _val(img, px, c, j, None, tz, kh, ty, n, e)
        assert abs(got - want) <= 2.0 ** -15 * abs(want) + 1e-30


def test_shortcut_and_two_pass_images():
    g = torch.Generator().manual_seed(5)
    ws = torch.randn((32, 64, 1, 1, 1), generator=g)
    img = pack_shortcut_weight_tc(ws, 2)
    assert tuple(img.shape) == (2, 4, 2, 2, 16, 8)            # [n-slice][Csrc/16][plane][khalf][n_cta][8]
    for sel, c, kh, n, e in ((0, 0, 0, 0, 0), (1, 3, 1, 15, 7), (1, 2, 0, 4, 5)):
        want = ws[sel * 16 + n, 16 * c + 8 * kh + e, 0, 0, 0].item()
        assert abs(_val(img, sel, c, None, kh, n, e) - want) <= 2.0 ** -15 * abs(want)
    # two-pass image of a Cout = 1 conv: plane 0 = [hi | lo | 0...], plane 1 = [hi | 0...]; exact bf16 values
    w1 = torch.randn((1, 16, 3, 3, 1), generator=g)
    tp = pack_conv_weight_tc_2p(w1)                           # [1][1][3][plane][1][2][3][16][8]
    hi = w1.bfloat16()
    lo = (w1 - hi.float()).bfloat16()
    for j, typ, kh, e in ((0, 0, 0, 0), (2, 1, 1, 3), (1, 2, 0, 7)):
        ci, ty = 8 * kh + e, 2 - typ
        assert tp[0, 0, j, 0, 0, kh, typ, 0, e] == hi[0, ci, j, ty, 0] and tp[0, 0, j, 0, 0, kh, typ, 1, e] == lo[0, ci, j, ty, 0]
        assert tp[0, 0, j, 1, 0, kh, typ, 0, e] == hi[0, ci, j, ty, 0] and tp[0, 0, j, 1, 0, kh, typ, 1, e] == 0
        assert tp[0, 0, j, 0, 0, kh, typ, 2:, e].abs().max() == 0


def test_folded_epilogue_equals_bias_batchnorm_prelu():
    torch.manual_seed(6)
    conv = torch.nn.Conv3d(4, 10, 1)
    bn = torch.nn.BatchNorm3d(10).eval()
    act = torch.nn.PReLU()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
        act.weight.fill_(0.17)
    sd = {"conv.bias": conv.bias.detach(), "norm.weight": bn.weight.detach(), "norm.bias": bn.bias.detach(),
          "norm.running_mean": bn.running_mean, "norm.running_var": bn.running_var, "act.weight": act.weight.detach()}
    scale, shift, slope, code = fold_epilogue(sd, "", 10, 16, True, "prelu")
    assert scale.shape == (16,) and (scale[10:] == 1).all() and (shift[10:] == 0).all() and code == 0
    x = torch.randn(2, 4, 3, 3, 3)
    with torch.no_grad():
        raw = torch.nn.functional.conv3d(x, conv.weight)                     # what the accumulator holds (no bias)
        want = act(bn(raw + conv.bias.view(1, -1, 1, 1, 1)))
        f = raw * scale[:10].view(1, -1, 1, 1, 1) + shift[:10].view(1, -1, 1, 1, 1)
        got = torch.where(f >= 0, f, f * slope)
    assert (got - want).abs().max().item() < 1e-5
    assert fold_epilogue(sd, "", 10, 16, False, "sigmoid")[2:] == (0.0, 1)
    s2, h2, sl2, _ = fold_epilogue(sd, "", 10, 16, False, "none")
    assert (s2 == 1).all() and torch.equal(h2[:10], conv.bias.detach().float()) and sl2 == 1.0

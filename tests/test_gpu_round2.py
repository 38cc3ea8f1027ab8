"""GPU parity tests added in round 2: the configuration bench.py times (128^3 windows, window groups, captured
graph, fused gate+logits launch), the device metric path, the native standalone DiceLoss and the 2-rank NCCL path.

Tolerances: 1e-3 max-abs on logits and identical argmax wherever the oracle margin exceeds 1e-4 (BASELINE.json);
block-level kernels 5e-5 of the activation scale (split-bf16 storage, fp32 accumulation)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import loss_oracle, sw_oracle, unet_oracle  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _native_net(sd, attention=True):
    from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA
    net = UNet2d5_spvPA(dimensions=3, in_channels=1, out_channels=2, channels=unet_oracle.CHANNELS,
                        strides=unet_oracle.STRIDES, kernel_sizes=unet_oracle.KERNEL_SIZES,
                        sample_kernel_sizes=unet_oracle.SAMPLE_KERNEL_SIZES, num_res_units=2, norm="BATCH",
                        dropout=0.1, attention_module=attention)
    net.load_state_dict(sd, strict=True)
    return net.to(_dev()).eval()


# ---- fused gate + logits launch -------------------------------------------------------------------------------
@pytest.mark.parametrize("B,dims,cout,gate,blend,multi", [
    (1, (6, 10, 16), 2, True, False, False),     # whole y range in one tile, no x segmentation effects
    (2, (9, 70, 8), 2, True, True, True),        # two y tiles (halo lines), one destination view per window
    (1, (40, 128, 32), 2, True, True, False),    # x segments + y tiles, blend into an accumulator region
    (2, (5, 12, 24), 1, False, False, False),    # no gate, single output channel
])
def test_gate_logits_matches_torch(B, dims, cout, gate, blend, multi):
    """vsseg_conv3d_gate_logits == conv3d(x * (1 + att), W, pad (1,1,0)) + bias [blended with the weight map]."""
    import ctypes as C
    from vs_seg_b200 import lib as vlib
    from vs_seg_b200.engine import pack_conv_weight
    from vs_seg_b200.tensors import Act8Buffer, f32view
    dev = _dev()
    lib = vlib.load()
    g = torch.Generator().manual_seed(B * 100 + dims[1])
    x = torch.randn((B, 32) + dims, generator=g)
    att = torch.rand((B, 1) + dims, generator=g)
    w = torch.randn((cout, 32, 3, 3, 1), generator=g) * 0.2
    bias = torch.randn(cout, generator=g)
    sw = torch.rand(dims, generator=g) + 0.1
    buf = Act8Buffer(B, 32, *dims, dev).from_ncdhw(x.to(dev))
    xq = buf.to_ncdhw().cpu()   # the split-bf16 rounded input both sides see
    ref = F.conv3d(xq * (1 + att) if gate else xq, w, bias, padding=(1, 1, 0))
    # destination: a region of a larger accumulator volume (strided view), pre-filled when blending
    big = torch.randn((B, cout, dims[0] + 3, dims[1] + 2, dims[2] + 8), generator=g)
    off = (2, 1, 8)
    big_d = big.to(dev)
    att_d, sw_d = att.to(dev), sw.to(dev)
    wh, bh = pack_conv_weight(w, False).contiguous(), bias.clone()
    xv = buf.view()
    av = f32view(att_d)
    if multi:
        arr = (vlib.F32View * B)()
        for b in range(B):
            C.memmove(C.byref(arr[b]), C.byref(f32view(big_d[b:b + 1], off, dims)), C.sizeof(vlib.F32View))
        outs, n_outs = arr, B
    else:
        ov = f32view(big_d, off, dims)
        outs, n_outs = C.byref(ov), 1
    vlib.check(lib.vsseg_conv3d_gate_logits(C.byref(xv), C.byref(av) if gate else None, wh.data_ptr(), bh.data_ptr(), cout,
                                            outs, n_outs, sw_d.data_ptr() if blend else None, 1 if multi else 0,
                                            torch.cuda.current_stream().cuda_stream), "gate_logits")
    got = big_d.cpu()
    region = (slice(None), slice(None)) + tuple(slice(o, o + d) for o, d in zip(off, dims))
    want = big.clone()
    want[region] = big[region] + sw * ref if blend else ref
    assert torch.equal(got[:, :, :2], big[:, :, :2])   # nothing outside the region is touched
    err = (got - want).abs().max().item()
    assert err < 5e-5 * max(1.0, ref.abs().max().item()), err


# ---- the timed configuration ------------------------------------------------------------------------------------
def test_window_group_plan_equals_single_window_plan_at_128():
    """Group-of-N plan == single-window plan, bit for bit, at the benchmark window size (line mode, x march and
    LZ = 128 flavours of the tensor-core kernel, the multi-destination gate+logits launch)."""
    from vs_seg_b200.tensors import f32view
    net = _native_net(unet_oracle.seeded_state_dict(0))
    roi = (128, 128, 128)
    vol = torch.randn((1, 1, 224, 128, 160), generator=torch.Generator().manual_seed(5)).to(_dev())
    starts = [(0, 0, 0), (96, 0, 32), (48, 0, 16)]
    imap = torch.rand(roi, generator=torch.Generator().manual_seed(6)).to(_dev())
    one = net.eval_plan(roi, batch=1)
    grp = net.eval_plan(roi, batch=len(starts), window_levels=1)
    acc1 = torch.zeros((1, 2, 224, 128, 160), device=_dev())
    acc2 = torch.zeros_like(acc1)
    for s_ in starts:
        one.run(f32view(vol, s_, roi), f32view(acc1, s_, roi), imap.data_ptr())
    grp.run([f32view(vol, s_, roi) for s_ in starts], [f32view(acc2, s_, roi) for s_ in starts], imap.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(acc1, acc2)
    assert not any(st.name.endswith(".gate") and st.name.startswith("dec0") for st in grp.steps)   # no top-level gate pass


@pytest.mark.parametrize("roi,vol_shape,starts", [
    ((128, 128, 128), (1, 1, 224, 128, 160), [(0, 0, 0), (96, 0, 32), (48, 0, 16)]),
    ((64, 64, 16), (2, 1, 96, 80, 24), [(0, 0, 0), (32, 16, 8), (16, 0, 4), (8, 16, 0), (24, 8, 8)]),
])
def test_batch_first_window_set_equals_per_window_launches(monkeypatch, roi, vol_shape, starts):
    """VSSEG_SW_BATCH_FIRST=1: the first ResidualUnit of every window in one launch per conv (the windows' source views
    travel as a window set, vsseg_f32view.n_windows) == one launch per window, bit for bit; the second case takes its
    windows from two batch entries of the volume."""
    from vs_seg_b200.tensors import f32view
    net = _native_net(unet_oracle.seeded_state_dict(0))
    vol = torch.randn(vol_shape, generator=torch.Generator().manual_seed(5)).to(_dev())
    imap = torch.rand(roi, generator=torch.Generator().manual_seed(6)).to(_dev())
    nb = vol_shape[0]
    accs = []
    for mode in ("0", "1"):
        monkeypatch.setenv("VSSEG_SW_BATCH_FIRST", mode)
        grp = net.eval_plan(roi, batch=len(starts), window_levels=1)
        names = [st.name for st in grp.steps]
        assert ("enc0.unit1" in names) == (mode == "1") and ("enc0.unit1@w0" in names) == (mode == "0")
        acc = torch.zeros((nb, 2) + tuple(vol_shape[2:]), device=_dev())
        srcs = [f32view(vol[i % nb:i % nb + 1], s_, roi) for i, s_ in enumerate(starts)]
        dsts = [f32view(acc[i % nb:i % nb + 1], s_, roi) for i, s_ in enumerate(starts)]
        for _ in range(2):   # the second run re-binds the same records
            acc.zero_()
            grp.run(srcs, dsts, imap.data_ptr())
        torch.cuda.synchronize()
        accs.append(acc)
    assert torch.equal(accs[0], accs[1])
    assert accs[0].abs().max().item() > 0


def test_atomic_out_plan_takes_all_windows_in_one_launch():
    """Plans built for an accumulator that several writers share (multi-GPU peer blend, vs_seg_b200/peer.py) blend
    every window of the group in ONE gate+logits launch with red.global.add; the sum order of overlapping windows is
    unspecified, so the result agrees with the per-window launches to fp32 rounding."""
    from vs_seg_b200.tensors import f32view
    net = _native_net(unet_oracle.seeded_state_dict(0))
    roi = (64, 64, 16)
    vol = torch.randn((1, 1, 96, 80, 24), generator=torch.Generator().manual_seed(8)).to(_dev())
    starts = [(0, 0, 0), (32, 16, 8), (16, 0, 4), (8, 16, 0)]
    imap = torch.rand(roi, generator=torch.Generator().manual_seed(9)).to(_dev())
    ref_plan = net.eval_plan(roi, batch=len(starts), window_levels=1)
    at_plan = net.eval_plan(roi, batch=len(starts), window_levels=1, atomic_out=True)
    assert sum(st.name.startswith("dec0.gate+logits") for st in at_plan.steps) == 1
    assert sum(st.name.startswith("dec0.gate+logits") for st in ref_plan.steps) == len(starts)
    acc1 = torch.zeros((1, 2, 96, 80, 24), device=_dev())
    acc2 = torch.zeros_like(acc1)
    srcs = [f32view(vol, s_, roi) for s_ in starts]
    ref_plan.run(srcs, [f32view(acc1, s_, roi) for s_ in starts], imap.data_ptr())
    at_plan.run(srcs, [f32view(acc2, s_, roi) for s_ in starts], imap.data_ptr(), atomic=True)
    torch.cuda.synchronize()
    assert (acc1 - acc2).abs().max().item() <= 1e-5 * max(1.0, acc1.abs().max().item())
    with pytest.raises(ValueError):
        at_plan.run(srcs, [f32view(acc2, s_, roi) for s_ in starts], imap.data_ptr())


def test_unet_eval_128_matches_oracle():
    """Whole-network eval forward at 128^3 (the benchmark window) vs the CPU oracle: logits and attention maps."""
    sd = unet_oracle.seeded_state_dict(0)
    net = _native_net(sd)
    x = torch.randn((1, 1, 128, 128, 128), generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        ref, ref_atts = unet_oracle.unet_forward(sd, x)
        got, atts = net(x.to(_dev()))
    got = got.cpu()
    assert (got - ref).abs().max().item() < 1e-3
    margin = (ref[:, 1] - ref[:, 0]).abs()
    assert ((got.argmax(1) != ref.argmax(1)) & (margin > 1e-4)).sum().item() == 0
    for a, r in zip(atts, ref_atts):
        assert a.shape == r.shape and (a.cpu() - r).abs().max().item() < 1e-4


def test_captured_graph_equals_direct_launches_and_follows_the_volume(monkeypatch):
    """The sliding-window program: (1) graph replay == direct launches bit for bit, (2) one captured graph serves
    another volume of the same layout (relocatable views), (3) ragged last group, (4) result vs the oracle."""
    from vs_seg_b200 import sliding_window as sw
    sd = unet_oracle.seeded_state_dict(4)
    net = _native_net(sd)
    roi = (64, 64, 16)
    g = torch.Generator().manual_seed(21)
    xa = torch.randn((1, 1, 96, 144, 24), generator=g)
    xb = torch.randn((1, 1, 96, 144, 24), generator=g)
    monkeypatch.setenv("VSSEG_SW_GROUP", "5")    # 2 x 3 x 2 = 12 windows -> groups of 5, 5, 2
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VSSEG_SW_GRAPH", mode)
        sw._PROGRAMS.clear()
        with torch.no_grad():
            outs[mode] = [sw.sliding_window_inference(t.to(_dev()), roi, 1, net, mode="gaussian").cpu() for t in (xa, xb, xa)]
        if mode == "1":
            assert len(sw._PROGRAMS) == 1 and next(iter(sw._PROGRAMS.values())).graph is not None
    for a, b in zip(outs["1"], outs["0"]):
        assert torch.equal(a, b)
    assert torch.equal(outs["1"][0], outs["1"][2]) and not torch.equal(outs["1"][0], outs["1"][1])
    with torch.no_grad():
        ref = sw_oracle.sliding_window_inference(xb, roi, 1, lambda w: unet_oracle.unet_forward(sd, w)[0], mode="gaussian")
    assert (outs["1"][1] - ref).abs().max().item() < 1e-3


def test_two_stream_window_groups_match_one_stream(monkeypatch):
    """VSSEG_SW_STREAMS=2 (two window groups with disjoint destinations run concurrently on two streams, joined per
    phase) == the one-stream schedule up to the fp32 order of overlapping windows of different phases."""
    from vs_seg_b200 import sliding_window as sw
    sd = unet_oracle.seeded_state_dict(4)
    net = _native_net(sd)
    roi = (64, 64, 16)
    x = torch.randn((1, 1, 160, 64, 16), generator=torch.Generator().manual_seed(23)).to(_dev())   # 4 x windows
    monkeypatch.setenv("VSSEG_SW_GROUP", "1")
    outs = []
    for streams in ("1", "2"):
        monkeypatch.setenv("VSSEG_SW_STREAMS", streams)
        sw._PROGRAMS.clear()
        with torch.no_grad():
            outs.append([sw.sliding_window_inference(x, roi, 1, net, mode="gaussian").clone() for _ in range(2)])
        prog = next(iter(sw._PROGRAMS.values()))
        assert prog.streams == int(streams)
        if streams == "2":
            assert any(len(ph) == 2 for ph in prog.phases)
    assert torch.equal(outs[1][0], outs[1][1])
    assert (outs[0][0] - outs[1][0]).abs().max().item() < 1e-5


def test_sliding_window_rejects_train_mode():
    from vs_seg_b200.sliding_window import sliding_window_inference
    net = _native_net(unet_oracle.seeded_state_dict(4)).train()
    with pytest.raises(RuntimeError):
        sliding_window_inference(torch.zeros((1, 1, 64, 64, 16), device=_dev()), (64, 64, 16), 1, net, mode="gaussian")


# ---- device metric path (VSparams.compute_dice_score, reference VSparams.py:393-408) ---------------------------
@pytest.mark.parametrize("shape,label_dtype", [((1, 2, 16, 16, 8), torch.float32), ((2, 2, 9, 7, 5), torch.float32),
                                                ((1, 2, 16, 16, 8), torch.uint8)])
def test_hard_dice_matches_oracle(shape, label_dtype):
    from vs_seg_b200.sliding_window import hard_dice
    g = torch.Generator().manual_seed(shape[2])
    p = torch.randn(shape, generator=g)
    label = (torch.rand((shape[0], 1) + shape[2:], generator=g) > 0.6).float()
    d, mask = hard_dice(p.to(_dev()), label.to(label_dtype).to(_dev()), return_mask=True)
    assert torch.equal(mask.cpu().long()[:, 0], p.argmax(1))
    for b in range(shape[0]):
        assert abs(d[b].item() - loss_oracle.dice_score(p[b:b + 1], label[b:b + 1]).item()) < 1e-6
    # empty label and empty prediction: (0 + eps) / (0 + eps) = 1
    z = torch.zeros((1, 1) + shape[2:])
    pz = torch.stack([torch.ones(shape[2:]), -torch.ones(shape[2:])])[None]
    assert abs(hard_dice(pz.to(_dev()), z.to(_dev())).item() - 1.0) < 1e-12


def test_vsparams_compute_dice_score_runs_on_the_device():
    from params.VSparams import VSparams
    p = VSparams.__new__(VSparams)
    p.device = _dev()
    g = torch.Generator().manual_seed(3)
    prob = torch.randn((1, 2, 32, 32, 8), generator=g)
    label = (torch.rand((1, 1, 32, 32, 8), generator=g) > 0.5).float()
    got = p.compute_dice_score(prob.to(_dev()), label.to(_dev()))
    assert got.shape == (1, 1) and got.is_cuda
    assert abs(got.item() - loss_oracle.dice_score(prob, label).item()) < 1e-6


def test_finalize_uint8_label_and_vector_tail():
    from vs_seg_b200.sliding_window import dice_from_sums, finalize
    dev = _dev()
    g = torch.Generator().manual_seed(2)
    for dims in ((16, 16, 8), (5, 7, 3)):   # n % 4 == 0 (128-bit path) and an odd size (scalar path)
        acc = torch.randn((1, 2) + dims, generator=g)
        cnt = torch.rand(dims, generator=g) + 0.5
        label = torch.rand((1, 1) + dims, generator=g) > 0.7
        out, mask, sums = finalize(acc.to(dev), cnt.to(dev), [0, 0, 0], list(dims), label=label.to(torch.uint8).to(dev),
                                   return_mask=True)
        ref = acc / cnt
        assert torch.equal(out.cpu(), ref)
        assert torch.equal(mask.cpu().long()[:, 0], ref.argmax(1))
        assert abs(dice_from_sums(sums)[0].item() - loss_oracle.dice_score(ref, label.float()).item()) < 1e-6


# ---- native standalone DiceLoss (reference dice_spvPA.py:90-167) -----------------------------------------------
DICE_FLAG_CASES = [
    dict(),                                                         # single channel, the attention-map form
    dict(to_onehot_y=True, softmax=True),                           # the logits form
    dict(to_onehot_y=True, softmax=True, include_background=False),
    dict(sigmoid=True, squared_pred=True),
    dict(to_onehot_y=True, softmax=True, jaccard=True, reduction="sum"),
    dict(to_onehot_y=True, softmax=True, reduction="none", weighted=True),
]


@pytest.mark.parametrize("flags", DICE_FLAG_CASES)
def test_dice_loss_native_matches_cpu_module(flags):
    """DiceLoss on CUDA tensors (native reduction + native backward) vs the same module on the CPU (the torch
    composition that follows the reference line by line and is pinned by tests/test_oracle_golden.py)."""
    from params.losses.dice_spvPA import DiceLoss
    flags = dict(flags)
    weighted = flags.pop("weighted", False)
    C = 2 if flags.get("softmax") or flags.get("to_onehot_y") else 1
    shape = (2, C, 12, 10, 8)
    g = torch.Generator().manual_seed(len(flags) + C)
    x = torch.randn(shape, generator=g)
    if C == 1 and not flags.get("sigmoid"):
        x = torch.rand(shape, generator=g)
    t = (torch.rand((2, 1) + shape[2:], generator=g) > 0.6).float()
    if not flags.get("to_onehot_y") and C > 1:
        t = loss_oracle.one_hot(t, C)
    w = torch.rand(shape, generator=g) + 0.5 if weighted else None
    xc = x.clone().requires_grad_(True)
    ref = DiceLoss(hardness_weight=w, **flags)(xc, t)
    ref.sum().backward()
    xg = x.to(_dev()).requires_grad_(True)
    got = DiceLoss(hardness_weight=w.to(_dev()) if w is not None else None, **flags)(xg, t.to(_dev()))
    got.sum().backward()
    assert got.shape == ref.shape
    assert (got.cpu() - ref).abs().max().item() < 2e-6
    assert (xg.grad.cpu() - xc.grad).abs().max().item() < 1e-6 + 1e-4 * xc.grad.abs().max().item()
    if C == 1 and not flags:
        assert abs(got.item() - loss_oracle.dice_loss(x, t).item()) < 2e-6


def test_dice_loss_native_shape_mismatch_raises():
    from params.losses.dice_spvPA import DiceLoss
    with pytest.raises(AssertionError):
        DiceLoss()(torch.zeros(1, 1, 4, 4, 4, device=_dev()), torch.zeros(1, 1, 4, 4, 5, device=_dev()))


def test_fused_adam_state_dict_carries_the_flat_moments():
    import copy
    from vs_seg_b200.optim import FusedAdam
    torch.manual_seed(0)
    ws = [torch.nn.Parameter(torch.randn(7, 3, device=_dev())), torch.nn.Parameter(torch.randn(5, device=_dev()))]
    vs = [torch.nn.Parameter(w.detach().clone()) for w in ws]
    a, b = FusedAdam(ws, lr=1e-2, weight_decay=1e-3), FusedAdam(vs, lr=1e-2, weight_decay=1e-3)
    g = [[torch.randn_like(w) for w in ws] for _ in range(4)]
    for i in range(2):
        a.zero_grad()
        for w, gi in zip(ws, g[i]):
            w.grad.copy_(gi)
        a.step()
    sd = copy.deepcopy(a.state_dict())
    with torch.no_grad():
        for v, w in zip(vs, ws):
            v.copy_(w)
    b.load_state_dict(sd)
    for i in range(2, 4):
        for opt, ps in ((a, ws), (b, vs)):
            opt.zero_grad()
            for w, gi in zip(ps, g[i]):
                w.grad.copy_(gi)
            opt.step()
    for v, w in zip(vs, ws):
        assert torch.equal(v.data, w.data)


def test_standalone_convolution_in_train_mode_with_grad_raises():
    from params.networks.blocks.convolutions import Convolution
    blk = Convolution(3, 16, 8, kernel_size=(3, 3, 1), act="RELU", norm=None, dropout=None).to(_dev()).train()
    x = torch.randn(1, 16, 8, 8, 8, device=_dev())
    with pytest.raises(NotImplementedError):
        blk(x)
    with torch.no_grad():
        assert blk(x).shape == (1, 8, 8, 8, 8)


@pytest.mark.parametrize("kind,cin,cout,k,stride,transposed,dims", [
    ("conv", 16, 32, (3, 3, 3), (1, 1, 1), False, (8, 8, 128)),
    ("conv", 32, 32, (3, 3, 1), (2, 2, 1), False, (8, 8, 128)),
    ("conv", 48, 32, (3, 3, 3), (2, 2, 2), True, (4, 4, 64)),
    ("ru", 32, 48, (3, 3, 3), (1, 1, 1), False, (8, 8, 128)),
])
def test_standalone_blocks_train_mode_match_torch_autograd(kind, cin, cout, k, stride, transposed, dims):
    """Train-mode standalone Convolution / ResidualUnit on CUDA (reference convolutions.py:148-156, :209-255 under
    autograd): batch-statistics BatchNorm, PReLU, running-stat update and every gradient vs the same module on the CPU
    (dropout 0: masks cannot be matched).  Tolerances as test_unet_train_step_matches_torch_autograd."""
    import copy
    from params.networks.blocks.convolutions import Convolution, ResidualUnit
    torch.manual_seed(cin + cout)
    if kind == "conv":
        ref = Convolution(3, cin, cout, strides=stride, kernel_size=k, act="PRELU", norm="BATCH", dropout=0.0,
                          is_transposed=transposed)
    else:
        ref = ResidualUnit(3, cin, cout, strides=1, kernel_size=k, subunits=2, act="PRELU", norm="BATCH", dropout=0.0)
    nat = copy.deepcopy(ref).to(_dev()).train()
    ref.train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, cin) + dims, generator=g)
    xr = x.clone().requires_grad_(True)
    xn = x.to(_dev()).requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    yn = nat(xn)
    yn.backward(gy.to(_dev()))
    assert yn.shape == yr.shape
    assert (yn.detach().cpu() - yr.detach()).abs().max().item() < 2e-3 * max(1.0, yr.abs().max().item())
    gmax_x = xr.grad.abs().max().item()
    assert (xn.grad.cpu() - xr.grad).abs().max().item() < 3e-2 * gmax_x
    pr, pn = dict(ref.named_parameters()), dict(nat.named_parameters())
    gmax_all = max(p.grad.abs().max().item() for p in pr.values())
    for name, p in pr.items():
        assert pn[name].grad is not None, name
        err = (pn[name].grad.cpu() - p.grad).abs().max().item()
        rel = 6e-2 if p.numel() == 1 else 3e-2
        assert err < rel * p.grad.abs().max().item() + 1e-4 * gmax_all, (name, err)
    for (n1, b1), (_, b2) in zip(ref.named_buffers(), nat.named_buffers()):
        assert (b2.cpu().float() - b1.float()).abs().max().item() < 1e-4 * max(1.0, b1.float().abs().max().item()), n1


def test_ts_mode_conv_parity_in_a_subprocess():
    """The TS flavour of the tensor-core kernel (A operand staged in tensor memory with tcgen05.cp, default off because it
    is slower on this network) stays correct: the conv / fused-shortcut / whole-net parity tests with VSSEG_TC_TS=2."""
    env = dict(os.environ, VSSEG_TC_TS="2")
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "--no-header",
           "-k", "tcgen05_conv or fused_shortcut or unet_eval_matches_reference_golden", "-p", "no:cacheprovider"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_graphed_train_step_matches_eager_steps():
    """The training step replayed as one CUDA graph (GraphedTrainStep) == the eager native step: same losses and the same
    parameters / BatchNorm statistics / Adam moments after several steps on changing batches (dropout 0; the weight
    gradients use fp32 atomics, so equality is to rounding), a learning-rate change in between is honoured, the capture's
    warm-up steps leave no trace, and tensor versions are bumped for the cached eval plans."""
    import copy
    from params.losses.dice_spvPA import Dice_spvPA
    from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA
    from vs_seg_b200.optim import FusedAdam
    from vs_seg_b200.training import GraphedTrainStep
    torch.manual_seed(0)
    a = UNet2d5_spvPA(dimensions=3, in_channels=1, out_channels=2, channels=unet_oracle.CHANNELS, strides=unet_oracle.STRIDES,
                      kernel_sizes=unet_oracle.KERNEL_SIZES, sample_kernel_sizes=unet_oracle.SAMPLE_KERNEL_SIZES,
                      num_res_units=2, norm="BATCH", dropout=0.0).to(_dev()).train()
    b = copy.deepcopy(a)
    oa, ob = FusedAdam(a.parameters(), lr=1e-4, weight_decay=1e-7), FusedAdam(b.parameters(), lr=1e-4, weight_decay=1e-7)
    crit = Dice_spvPA(to_onehot_y=True, softmax=True)
    graphed = GraphedTrainStep(b, crit, ob)
    g = torch.Generator().manual_seed(9)
    # one eager step on both models first: the graphed step must cope with autograd state left behind by eager steps
    x0 = torch.randn((2, 1, 64, 64, 16), generator=g).to(_dev())
    y0 = (torch.rand((2, 1, 64, 64, 16), generator=g) > 0.7).float().to(_dev())
    for m, opt in ((a, oa), (b, ob)):
        opt.zero_grad()
        crit(m(x0), y0).backward()
        opt.step()
    v0 = next(b.parameters())._version
    for i in range(4):
        x = torch.randn((2, 1, 64, 64, 16), generator=g).to(_dev())
        y = (torch.rand((2, 1, 64, 64, 16), generator=g) > 0.7).float().to(_dev())
        if i == 2:
            for opt in (oa, ob):
                opt.param_groups[0]["lr"] = 2.5e-5
        oa.zero_grad()
        la = crit(a(x), y)
        la.backward()
        oa.step()
        lb = graphed(x, y)
        assert abs(la.item() - lb.item()) < 1e-4 * max(1.0, abs(la.item())), (i, la.item(), lb.item())
    assert oa.param_groups[0]["step"] == ob.param_groups[0]["step"] == 5
    assert next(b.parameters())._version > v0
    # Adam normalises the update to ~lr per step whatever the gradient's size, so an element whose gradient is at the
    # rounding level of the atomics may move differently: compare the bulk tightly and bound the stragglers by the
    # total step length (4 steps x lr)
    d = torch.cat([(p1 - p2).detach().abs().reshape(-1) for p1, p2 in zip(a.parameters(), b.parameters())])
    assert (d > 1e-5).float().mean().item() < 0.01, (d > 1e-5).float().mean().item()
    assert d.max().item() < 8e-4, d.max().item()
    # the running statistics of the deep levels see the accumulated parameter differences above
    for (n1, b1), (_, b2) in zip(a.named_buffers(), b.named_buffers()):
        assert (b1.float() - b2.float()).abs().max().item() < 5e-3 * max(1.0, b1.float().abs().max().item()), n1


def test_train_mode_dropout_mask_statistics_and_backward_consistency():
    """Dropout inside the native train-mode block (reference convolutions.py:148-156: Conv -> BN -> Dropout -> PReLU):
    the keep fraction is 1 - p, kept values are scaled by 1/(1-p), and the backward pass regenerates exactly the mask of
    the forward pass - checked by replaying the block in torch with the mask read off the native output."""
    import copy
    from params.networks.blocks.convolutions import Convolution
    p_drop = 0.25
    torch.manual_seed(7)
    blk = Convolution(3, 16, 32, strides=1, kernel_size=(3, 3, 3), act="PRELU", norm="BATCH", dropout=p_drop)
    ref = copy.deepcopy(blk).train()
    nat = blk.to(_dev()).train()
    g = torch.Generator().manual_seed(8)
    x = torch.randn((2, 16, 8, 8, 128), generator=g)
    gy = torch.randn((2, 32, 8, 8, 128), generator=g)
    xn = x.to(_dev()).requires_grad_(True)
    yn = nat(xn)
    yn.backward(gy.to(_dev()))
    y = yn.detach().cpu()
    keep = (y != 0)
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p_drop)) < 5e-3, frac
    # torch replay with the same mask
    xr = x.clone().requires_grad_(True)
    u = ref.norm(ref.conv(xr))
    yr = torch.nn.functional.prelu(u * keep.float() / (1 - p_drop), ref.act.weight)
    yr.backward(gy)
    assert (y - yr.detach()).abs().max().item() < 2e-3 * max(1.0, yr.abs().max().item())
    assert (xn.grad.cpu() - xr.grad).abs().max().item() < 3e-2 * xr.grad.abs().max().item()
    gmax_all = max(p1.grad.abs().max().item() for p1 in ref.parameters())
    for (n1, p1), (_, p2) in zip(ref.named_parameters(), nat.named_parameters()):
        err = (p2.grad.cpu() - p1.grad).abs().max().item()
        # (the conv bias cancels in train-mode BatchNorm: its gradient is rounding noise on both sides)
        assert err < (6e-2 if p1.numel() == 1 else 3e-2) * p1.grad.abs().max().item() + 1e-4 * gmax_all, (n1, err)


# ---- two ranks over NCCL ----------------------------------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_rank_nccl_sharded_inference_equals_one_gpu(tmp_path):
    """Windows sharded over 2 NCCL ranks + reduce == the 1-GPU result (sum order differs: 1e-5) and the oracle."""
    out = str(tmp_path / "r0.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "nccl_sw_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = torch.load(out)
    assert res["err_vs_single"] < 1e-5, res
    assert res["err_vs_oracle"] < 1e-3 and res["flips"] == 0, res
    assert res["mask_equal"] and abs(res["dice_sharded"] - res["dice_single"]) < 1e-4, res

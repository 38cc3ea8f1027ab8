"""Host-side launch planner of the tensor-core conv kernel (vsseg_conv3d_tc_describe runs without a GPU): every conv
of UNet2d5_spvPA (reference unet2d5_spvPA.py:106-202, per-convolution table in SURVEY.md §8a) is covered at the
benchmark window and at awkward crops, the measured tile hints and the forced tile are honoured, and a decoder unit's
shortcut (convolutions.py:241-255, same input as the conv) is planned without shortcut stages."""
import ctypes as C
import re

import pytest

from vs_seg_b200 import lib as L

CH = (16, 32, 48, 64, 80, 96)
STRIDES = ((2, 2, 1), (2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2))
KS = ((3, 3, 1), (3, 3, 1), (3, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))


def _act(B, Cc, dims, ptr=4096):
    n = dims[0] * dims[1] * dims[2]
    return L.Act8(ptr, B * Cc * n, Cc * n, B, Cc, *dims)


def _describe(B, cin, cout, dims, k, s=(1, 1, 1), tr=False, sc=None, sc_ptr=4096):
    lib = L.load()
    od = tuple(d * q for d, q in zip(dims, s)) if tr else tuple((d + q - 1) // q for d, q in zip(dims, s))
    a, o = _act(B, cin, dims), _act(B, cout, od, ptr=8192)
    g = L.ConvGeom(*k, *s, 1 if tr else 0)
    scv = _act(B, sc, dims, ptr=sc_ptr) if sc else None
    scp = C.byref(scv) if scv is not None else None
    ns = lib.vsseg_conv3d_tc_suggest_split(C.byref(a), C.byref(o), C.byref(g), scp)
    if ns <= 0:
        return 0, {}
    buf = C.create_string_buffer(65536)
    assert lib.vsseg_conv3d_tc_describe(C.byref(a), C.byref(o), C.byref(g), ns, scp, buf, 65536) == 0
    head, _, ops = buf.value.decode().partition(" ops:")
    d = {k_: v for k_, v in re.findall(r"(\w+)=(\S+)", head)}
    main, _, sc_ops = ops.partition(" ops2:")
    op = lambda t: [tuple(int(v) for v in m) for m in re.findall(r"\(a(\d+) b(\d+) c(\d+) n(\d+)\)", t)]  # noqa: E731
    d["ops"], d["ops2"] = op(main), op(sc_ops)
    return ns, d


def _layers(patch):
    """(name, cin, cout, input dims, kernel, stride, transposed, shortcut channels, shortcut reads the conv input)."""
    dims = [tuple(patch)]
    for s in STRIDES:
        dims.append(tuple(d // q for d, q in zip(dims[-1], s)))
    out = []
    for l in range(5):
        if l > 0:
            out.append((f"enc{l}.unit0", CH[l - 1], CH[l], dims[l], KS[l], (1, 1, 1), False, None, False))
        out.append((f"enc{l}.unit1", CH[l], CH[l], dims[l], KS[l], (1, 1, 1), False, CH[l - 1] if l else None, False))
        out.append((f"down{l}", CH[l], CH[l], dims[l], KS[l], STRIDES[l], False, None, False))
        out.append((f"up{l}", CH[l + 1], CH[l], dims[l + 1], KS[l], STRIDES[l], True, None, False))
        out.append((f"dec{l}.att.conv1", 2 * CH[l], CH[l], dims[l], KS[l], (1, 1, 1), False, None, False))
        if l > 0:
            out.append((f"dec{l}.unit0", 2 * CH[l], CH[l], dims[l], KS[l], (1, 1, 1), False, 2 * CH[l], True))
    out.append(("bottom.unit0", CH[4], CH[5], dims[5], KS[5], (1, 1, 1), False, None, False))
    out.append(("bottom.unit1", CH[5], CH[5], dims[5], KS[5], (1, 1, 1), False, CH[4], False))
    return out


@pytest.mark.parametrize("patch,B", [((128, 128, 128), 8), ((128, 128, 128), 1), ((128, 128, 32), 2), ((96, 64, 40), 1),
                                     ((384, 384, 64), 1), ((64, 64, 16), 4)])
def test_every_conv_of_the_network_has_a_tensor_core_plan(patch, B):
    for name, cin, cout, dims, k, s, tr, sc, same in _layers(patch):
        if cin % 16:
            continue   # Cin = 1 first conv: conv_cin1_k331_kernel
        ns, d = _describe(B, cin, cout, dims, k, s, tr, sc, 4096 if same else 12288)
        assert ns >= 1, f"{name} at {patch}: no tcgen05 plan"
        assert int(d["tmem_cols"]) <= 512 and int(d["smem"]) <= 227 * 1024 and int(d["nstage"]) >= 2, (name, d)
        assert int(d["sc_self"]) == (1 if same else 0), (name, d)
        assert int(d["nop2"]) == (3 * int(d["YL"]) // int(d["LY"]) if sc else 0), (name, d)
        assert len(d["ops"]) == int(d["nop"]) and len(d["ops2"]) == int(d["nop2"])
        if same:
            # the shortcut's A views are views the conv itself reads from the same stage (its centre tap), one per
            # line group, hi*hi / lo*hi / hi*lo like the main products, into accumulator columns y * n_cta
            views = {a_ for a_, _, _, _ in d["ops"]}
            n_cta = int(d["n_cta"])
            for i, (a_, b_, c_, n_) in enumerate(d["ops2"]):
                assert a_ in views and n_ == n_cta and c_ == (i // 3) * n_cta, (name, i, d["ops2"])
            assert [b_ for _, b_, _, _ in d["ops2"][:3]] == [0, 0, n_cta * 2], (name, d["ops2"][:3])


def test_measured_tile_hints_apply_to_window_groups_only():
    """kTileHints (vsseg_tc.cu; profiles/r02_autotune_tiles*.tsv) are keyed by the layer geometry and used for batches
    of at least four windows; smaller batches go through the cost model."""
    _, up2 = _describe(8, 64, 48, (16, 16, 64), (3, 3, 3), (2, 2, 2), True)
    assert (up2["XT"], up2["YL"], up2["LY"]) == ("1", "2", "2")          # YT = 1 line group of LY = 2 lines
    _, enc = _describe(8, 32, 32, (64, 64, 128), (3, 3, 1), sc=16, sc_ptr=12288)
    assert (enc["XT"], enc["YL"]) == ("1", "4")
    _, enc1 = _describe(1, 32, 32, (64, 64, 128), (3, 3, 1), sc=16, sc_ptr=12288)
    assert enc1["XT"] != "1"                                             # the model's x march for a single window


def test_forced_tile_is_honoured_and_infeasible_tiles_are_refused(monkeypatch):
    monkeypatch.setenv("VSSEG_TC_FORCE", "4,2,3")
    ns, d = _describe(8, 64, 32, (64, 64, 128), (3, 3, 1))
    assert ns == 1 and (d["XT"], d["YL"], d["nstage"]) == ("4", "2", "3")
    monkeypatch.setenv("VSSEG_TC_FORCE", "1,64,0")                       # 64 lines x 32 columns do not fit 512 TMEM columns
    ns, _ = _describe(8, 64, 32, (64, 64, 128), (3, 3, 1))
    assert ns == 0

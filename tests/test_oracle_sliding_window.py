"""Pins of the sliding-window restatement (SURVEY.md §8c iii-iv; reference call site
/root/reference/params/VSparams.py:568-574).  MONAI 0.4.0 itself is not installable offline, so the pins are:
  (1) oracle/sw_oracle.py == tests/golden/sw_geometry.json (frozen by oracle/make_sw_golden.py);
  (2) the product's host logic (vs_seg_b200.sliding_window) == the same file;
  (3) window lists worked by hand from the published algorithm, the closed-form erf weights, and the
      SURVEY.md §8a S1 probe numbers (32 / 1 / 8 / 12 windows, w[0] = 3.3592e-4 w[64], 3-D minimum
      3.8e-11, multiplicity histogram 4,194,304 / 10,485,760 / 7,340,032 / 1,572,864).
"""
import itertools
import json
import math
import os

import pytest
import torch

from oracle import sw_oracle
from vs_seg_b200 import sliding_window as sw


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "sw_geometry.json")))


# window starts per axis, derived by hand: interval = int(roi * 0.75) unless roi == image; the last window
# snaps back so it ends at the image border
HAND = {
    ((384, 384, 160), (128, 128, 128)): ([0, 96, 192, 256], [0, 96, 192, 256], [0, 32]),
    ((384, 384, 64), (384, 384, 64)): ([0], [0], [0]),
    ((448, 448, 80), (384, 384, 64)): ([0, 64], [0, 64], [0, 16]),
    ((512, 512, 120), (384, 384, 64)): ([0, 128], [0, 128], [0, 48, 56]),
    ((128, 128, 64), (128, 128, 32)): ([0], [0], [0, 24, 32]),
    # image smaller than the roi in x and z: padded to (128, 140, 32) first (14 / 6 voxels before)
    ((100, 140, 20), (128, 128, 32)): ([0], [0, 12], [0]),
    ((96, 80, 24), (64, 64, 16)): ([0, 32], [0, 16], [0, 8]),
}


def test_window_lists_oracle_product_golden_and_hand(golden):
    assert len(golden["geometries"]) == len(HAND)
    for e in golden["geometries"]:
        image, roi = tuple(e["image"]), tuple(e["roi"])
        padded = tuple(max(i, r) for i, r in zip(image, roi))
        want = [tuple(s) for s in e["starts"]]
        # first spatial dim slowest, last fastest
        assert want == list(itertools.product(*HAND[(image, roi)]))
        interval = sw_oracle.scan_interval(padded, roi, 0.25)
        assert list(interval) == e["interval"]
        assert sw_oracle.window_starts(padded, roi, interval) == want
        assert sw.window_starts(padded, roi, 0.25) == want
        x = torch.zeros((1, 1) + image)
        _, lows = sw._pad_to_roi(x, roi, "constant", 0.0)
        assert lows == e["pad_low"]
    counts = {tuple(e["image"]): len(e["starts"]) for e in golden["geometries"] if tuple(e["roi"]) != (64, 64, 16)}
    assert counts[(384, 384, 160)] == 32 and counts[(384, 384, 64)] == 1
    assert counts[(448, 448, 80)] == 8 and counts[(512, 512, 120)] == 12


def _closed_form_axis(n, sigma_scale=0.125):
    sigma = n * sigma_scale
    tail = int(max(sigma * 4.0, 0.5) + 0.5)
    t = 0.70710678 / sigma
    out = []
    for i in range(n):
        d = i - n // 2
        w = 0.5 * (math.erf(t * (d + 0.5)) - math.erf(t * (d - 0.5)))
        out.append(max(w, 0.0) if abs(d) <= tail else 0.0)
    return out


def test_importance_map_oracle_product_golden_and_closed_form(golden):
    for e in golden["importance"]:
        roi = tuple(e["roi"])
        mo = sw_oracle.importance_map(roi, "gaussian", 0.125)
        mp = sw.importance_map(roi, "gaussian", 0.125, "cpu")
        assert mo.dtype == torch.float32 and mo.shape == roi
        # the product builds the same separable map by outer products: equal up to fp32 rounding of the order
        assert torch.allclose(mo, mp, rtol=2e-6, atol=0)
        for m in (mo.double(), mp.double()):
            assert m.max().item() == 1.0
            assert abs(m.min().item() / e["min"] - 1) < 1e-5
            assert abs(m[0, 0, 0].item() / e["corner"] - 1) < 1e-5
            assert abs(m.sum().item() / e["sum"] - 1) < 1e-6
            for d in range(3):
                idx = [r // 2 for r in roi]
                idx[d] = slice(None)
                prof = m[tuple(idx)]
                assert torch.allclose(prof, torch.tensor(e["axis_profiles"][d], dtype=torch.float64), rtol=1e-5, atol=0)
                cf = torch.tensor(_closed_form_axis(roi[d]), dtype=torch.float64)
                cf = cf / cf.max()
                # MONAI evaluates erf in fp32: the difference of two erf values near +-1 carries an absolute
                # error of a few fp32 ulps of 1, divided by the small centre weight (roi 128: the 3.3592e-4 tail is
                # 3.3634e-4 in exact arithmetic)
                assert torch.allclose(prof, cf, rtol=0, atol=5e-6)   # profiles are normalised to max 1
    # SURVEY.md §8a S1 probe numbers for roi 128: per-axis w[0] = 3.3592e-4 * w[64], 3-D minimum 3.8e-11 (never
    # zero, so the clamp-to-min-nonzero is a no-op)
    e = [x for x in golden["importance"] if tuple(x["roi"]) == (128, 128, 128)][0]
    a = e["axis_profiles"][0]
    assert abs(a[0] / a[64] - 3.3592e-4) < 5e-9
    assert abs(e["min"] - 3.8e-11) < 0.05e-11 and e["min"] > 0
    assert abs(e["min"] - (a[0] / a[64]) ** 3) / e["min"] < 1e-5


def test_multiplicity_histogram_of_the_benchmark_geometry():
    """How many windows cover each voxel of 384x384x160 with a 128^3 window."""
    image, roi = (384, 384, 160), (128, 128, 128)
    cnt = torch.zeros(image, dtype=torch.int32)
    for s in sw.window_starts(image, roi, 0.25):
        cnt[s[0]:s[0] + 128, s[1]:s[1] + 128, s[2]:s[2] + 128] += 1
    hist = {int(k): int((cnt == k).sum()) for k in cnt.unique()}
    assert hist == {1: 4194304, 2: 10485760, 4: 7340032, 8: 1572864}


def test_oracle_blend_matches_product_generic_path_and_padding():
    """End to end on CPU with an analytic predictor: the product's generic path == the oracle, for a
    volume smaller than the roi in one axis (constant padding, crop of the result) and sw_batch_size > 1."""
    def predictor(w):   # two 'classes' that depend on the window content and on the position inside the window
        ramp = torch.linspace(0, 1, w.shape[-1]).view(1, 1, 1, 1, -1)
        return torch.cat([w * 2 + ramp, -w + 0.5], 1)

    g = torch.Generator().manual_seed(7)
    for image, roi in (((40, 72, 12), (32, 32, 16)), ((50, 33, 20), (32, 32, 16))):
        x = torch.randn((2, 1) + image, generator=g)
        ref = sw_oracle.sliding_window_inference(x, roi, 3, predictor, mode="gaussian")
        got = sw.sliding_window_inference(x, roi, 3, predictor, mode="gaussian")
        assert got.shape == ref.shape == (2, 2) + image
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
        ref_c = sw_oracle.sliding_window_inference(x, roi, 1, predictor, mode="constant")
        got_c = sw.sliding_window_inference(x, roi, 1, predictor, mode="constant")
        assert torch.allclose(got_c, ref_c, rtol=1e-5, atol=1e-6)


def test_two_stream_phases_pair_only_groups_with_disjoint_destinations():
    """The sliding-window program runs two window groups at a time on two streams with a plain (non-atomic) blend
    (vs_seg_b200/sliding_window.py: _Program._pair_groups): the two groups of a phase must never touch the same
    accumulator voxel, every group runs exactly once, and unpaired groups stay alone."""
    import numpy as np
    for image, roi, group, batch in (((384, 384, 160), (128, 128, 128), 8, 1), ((448, 448, 80), (384, 384, 64), 2, 1),
                                     ((96, 144, 24), (64, 64, 16), 5, 2), ((160, 64, 16), (64, 64, 16), 1, 1),
                                     ((128, 128, 64), (128, 128, 32), 3, 1)):
        starts = sw.window_starts(image, roi, 0.25)
        jobs = [(b, s) for b in range(batch) for s in starts]
        groups = [jobs[g0:g0 + group] for g0 in range(0, len(jobs), group)]
        phases = sw._Program._pair_groups(groups, roi)
        assert sorted(i for ph in phases for i in ph) == list(range(len(groups)))
        assert all(1 <= len(ph) <= 2 for ph in phases)
        for ph in phases:
            if len(ph) < 2:
                continue
            marks = []
            for gi in ph:
                m = np.zeros((batch,) + tuple(image), dtype=bool)
                for b, s in groups[gi]:
                    m[b, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] = True
                marks.append(m)
            assert not (marks[0] & marks[1]).any(), (image, roi, ph)
    # the benchmark geometry pairs its four x slabs two by two
    starts = sw.window_starts((384, 384, 160), (128, 128, 128), 0.25)
    groups = [[(0, s) for s in starts[g0:g0 + 8]] for g0 in range(0, 32, 8)]
    assert sw._Program._pair_groups(groups, (128, 128, 128)) == [[0, 2], [1, 3]]

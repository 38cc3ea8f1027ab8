"""Generate tests/golden/sw_geometry.json from the sliding-window oracle (SURVEY.md §8c iii-iv).

Run:  python oracle/make_sw_golden.py

MONAI 0.4.0 is not installable offline, so these vectors cannot come from MONAI itself ("parity
unpinned" against MONAI, see oracle/sw_oracle.py).  They freeze the restatement instead, and
tests/test_oracle_sliding_window.py checks them three ways: (1) oracle == this file, (2) the product's
host logic (vs_seg_b200.sliding_window) == this file, (3) values worked by hand / in closed form from the
published algorithm and the SURVEY.md §8a S1 probe numbers (window lists, w[0]/w[64] = 3.3592e-4, the
multiplicity histogram of the 384x384x160 benchmark geometry).  Reference call site:
/root/reference/params/VSparams.py:568-574.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import sw_oracle  # noqa: E402

# (image, roi): the benchmark geometry, the reference default roi on three volume sizes, the --debug crop,
# an image smaller than the roi (padded first) and a ragged one
GEOMETRIES = [
    ((384, 384, 160), (128, 128, 128)),
    ((384, 384, 64), (384, 384, 64)),
    ((448, 448, 80), (384, 384, 64)),
    ((512, 512, 120), (384, 384, 64)),
    ((128, 128, 64), (128, 128, 32)),
    ((100, 140, 20), (128, 128, 32)),
    ((96, 80, 24), (64, 64, 16)),
]
IMAP_ROIS = [(128, 128, 128), (384, 384, 64), (128, 128, 32), (64, 64, 16)]


def main():
    out = {"geometries": [], "importance": []}
    for image, roi in GEOMETRIES:
        padded = tuple(max(i, r) for i, r in zip(image, roi))
        interval = sw_oracle.scan_interval(padded, roi, 0.25)
        starts = sw_oracle.window_starts(padded, roi, interval)
        pad_lo = [max(r - i, 0) // 2 for i, r in zip(image, roi)]
        out["geometries"].append({"image": image, "roi": roi, "padded": padded, "interval": interval,
                                  "pad_low": pad_lo, "starts": starts})
    for roi in IMAP_ROIS:
        m = sw_oracle.importance_map(roi, "gaussian", 0.125).double()
        axes = []
        for d, r in enumerate(roi):   # per-axis profile through the centre
            idx = [x // 2 for x in roi]
            idx[d] = slice(None)
            axes.append([float(v) for v in m[tuple(idx)]])
        out["importance"].append({
            "roi": roi, "sigma_scale": 0.125, "axis_profiles": axes,
            "min": float(m.min()), "max": float(m.max()), "sum": float(m.sum()),
            "corner": float(m[0, 0, 0]), "centre": float(m[tuple(r // 2 for r in roi)]),
        })
    path = os.path.join(ROOT, "tests", "golden", "sw_geometry.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

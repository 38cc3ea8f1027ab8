"""Per-launch device time of one 128^3 patch forward (CUDA events), written as a TSV.
Usage (GPU box): python tools/profile_plan.py [out.tsv]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vs_seg_b200 import sliding_window as sw  # noqa: E402
from vs_seg_b200.tensors import f32view  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "plan_profile.tsv")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    dev = torch.device("cuda:0")
    net, _ = bench.build_net(dev)
    roi = bench.ROI
    group = int(os.environ.get("PROFILE_GROUP", "1"))   # > 1: a window-group plan (per-patch = total / group)
    vol = torch.randn((1, 1) + bench.VOLUME, device=dev)
    acc = torch.zeros((1, 2) + bench.VOLUME, device=dev)
    imap = sw.importance_map(roi, "gaussian", 0.125, dev)
    if group > 1:
        plan = net.eval_plan(roi, group, dev, window_levels=int(os.environ.get("VSSEG_SW_WINDOW_LEVELS", "1")))
        starts = sw.window_starts(bench.VOLUME, roi, 0.25)[:group]
        prof = plan.profile([f32view(vol, s, roi) for s in starts], [f32view(acc, s, roi) for s in starts],
                            imap.data_ptr(), iters=5)
    else:
        plan = net.eval_plan(roi, 1, dev)
        prof = plan.profile(f32view(vol, (0, 0, 0), roi), f32view(acc, (0, 0, 0), roi), imap.data_ptr(), iters=5)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    tot = sum(p[4] for p in prof)
    lines = ["name\tkind\tGFLOP\tMB\tms\tshare\tTFLOP/s\tGB/s\troof_ms\tfrac_of_roof"]
    roof_tot = 0.0
    for name, kind, fl, nb, ms in prof:
        roof = max(fl / (peaks["bf16_tflops"] * 1e12), nb / (peaks["hbm_gbs"] * 1e9)) * 1e3
        roof_tot += roof
        lines.append(f"{name}\t{kind}\t{fl / 1e9:.3f}\t{nb / 1e6:.2f}\t{ms:.4f}\t{ms / tot:.3f}\t"
                     f"{fl / ms / 1e9:.2f}\t{nb / ms / 1e6:.1f}\t{roof:.4f}\t{roof / ms:.3f}")
    lines.append(f"TOTAL\t-\t{sum(p[2] for p in prof) / 1e9:.2f}\t{sum(p[3] for p in prof) / 1e6:.1f}\t{tot:.3f}\t1\t"
                 f"{sum(p[2] for p in prof) / tot / 1e9:.2f}\t{sum(p[3] for p in prof) / tot / 1e6:.1f}\t"
                 f"{roof_tot:.4f}\t{roof_tot / tot:.3f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "--host"):
    main()


def host_cost():
    """Host-side cost of issuing one patch (no device sync inside the loop)."""
    import time
    dev = torch.device("cuda:0")
    net, _ = bench.build_net(dev)
    plan = net.eval_plan(bench.ROI, 1, dev)
    vol = torch.randn((1, 1) + bench.VOLUME, device=dev)
    acc = torch.zeros((1, 2) + bench.VOLUME, device=dev)
    imap = sw.importance_map(bench.ROI, "gaussian", 0.125, dev)
    a, b = f32view(vol, (0, 0, 0), bench.ROI), f32view(acc, (0, 0, 0), bench.ROI)
    for _ in range(3):
        plan.run(a, b, imap.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        plan.run(a, b, imap.data_ptr())
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host issue time per patch {1e3 * (t1 - t0) / n:.3f} ms; wall incl. drain {1e3 * (t2 - t0) / n:.3f} ms; "
          f"{len(plan.steps)} launches")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "--host":
    host_cost()

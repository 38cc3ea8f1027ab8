"""Per-layer tile sweep of the tensor-core conv launches of one window group (GPU box):
    PROFILE_GROUP=8 python tools/autotune_tiles.py [out.tsv]
Every tcgen05 launch of the group plan is re-timed with the tile forced through VSSEG_TC_FORCE="XT,YT[,nstage]"
(x rows per tile, y line groups per tile, ring depth); the planner's own choice is the row 'default'.  Timing only:
every tile shape runs the same kernel code, the results are checked by the parity tests once a choice is adopted."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vs_seg_b200 import lib as L  # noqa: E402
from vs_seg_b200 import sliding_window as sw  # noqa: E402
from vs_seg_b200.tensors import f32view  # noqa: E402


def time_step(st, s, stream, iters):
    if st.fn(*st.args, s):
        return None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(stream)
        for _ in range(iters):
            st.fn(*st.args, s)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "autotune_tiles.tsv")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    dev = torch.device("cuda:0")
    net, _ = bench.build_net(dev)
    roi = bench.ROI
    group = int(os.environ.get("PROFILE_GROUP", "8"))
    vol = torch.randn((1, 1) + bench.VOLUME, device=dev)
    acc = torch.zeros((1, 2) + bench.VOLUME, device=dev)
    imap = sw.importance_map(roi, "gaussian", 0.125, dev)
    plan = net.eval_plan(roi, group, dev, window_levels=1)
    starts = sw.window_starts(bench.VOLUME, roi, 0.25)[:group]
    plan._bind([f32view(vol, s_, roi) for s_ in starts], [f32view(acc, s_, roi) for s_ in starts], imap.data_ptr())
    stream = torch.cuda.current_stream(dev)
    s = stream.cuda_stream
    for st in plan.steps:
        L.check(st.fn(*st.args, s), st.name)
    torch.cuda.synchronize()
    only = os.environ.get("AUTOTUNE_ONLY", "")
    xts = [int(v) for v in os.environ.get("AUTOTUNE_XT", "1,2,4,8,16,32,64").split(",")]
    yts = [int(v) for v in os.environ.get("AUTOTUNE_YT", "1,2,4,8,16").split(",")]
    nsts = [int(v) for v in os.environ.get("AUTOTUNE_NST", "0").split(",")]
    lines = ["layer\ttile\tms"]
    for st in plan.steps:
        if st.kind != "tcgen05" or "@w" in st.name and not st.name.endswith("@w0"):
            continue
        if only and not any(o in st.name for o in only.split(",")):
            continue
        os.environ.pop("VSSEG_TC_FORCE", None)
        base = time_step(st, s, stream, 5)
        rows = [("default", base)]
        for xt in xts:
            for yt in yts:
                for nst in nsts:
                    os.environ["VSSEG_TC_FORCE"] = f"{xt},{yt},{nst}"
                    t = time_step(st, s, stream, 5)
                    if t is not None:
                        rows.append((f"{xt},{yt},{nst}", t))
        os.environ.pop("VSSEG_TC_FORCE", None)
        rows.sort(key=lambda r: r[1])
        best = rows[0]
        print(f"{st.name}: default {base:.4f} ms; best {best[0]} {best[1]:.4f} ms ({100 * (1 - best[1] / base):.1f} % faster); "
              + " ".join(f"[{n} {t:.4f}]" for n, t in rows[:6]), flush=True)
        lines += [f"{st.name}\t{n}\t{t:.4f}" for n, t in rows]
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

"""Runs one window-group forward inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off) and
writes the launch order (step names) next to it:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/plan \
        python tools/ncu_plan.py gpurun_out/plan_steps.json
    ncu -i gpurun_out/plan.ncu-rep --page raw --csv > gpurun_out/plan_raw.csv
    python tools/ncu_summarise.py gpurun_out/plan_raw.csv gpurun_out/plan_steps.json profiles/r01_ncu_group_summary.json"""
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
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "plan_steps.json")
    dev = torch.device("cuda:0")
    net, _ = bench.build_net(dev)
    group = int(os.environ.get("PROFILE_GROUP", "8"))
    roi = bench.ROI
    vol = torch.randn((1, 1) + bench.VOLUME, device=dev)
    acc = torch.zeros((1, 2) + bench.VOLUME, device=dev)
    imap = sw.importance_map(roi, "gaussian", 0.125, dev)
    plan = net.eval_plan(roi, group, dev, window_levels=int(os.environ.get("VSSEG_SW_WINDOW_LEVELS", "1")))
    starts = sw.window_starts(bench.VOLUME, roi, 0.25)[:group]
    srcs, dsts = [f32view(vol, s, roi) for s in starts], [f32view(acc, s, roi) for s in starts]
    for _ in range(2):
        plan.run(srcs, dsts, imap.data_ptr())
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.run(srcs, dsts, imap.data_ptr())
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    json.dump({"group": group, "steps": [{"name": st.name, "kind": st.kind, "flops": st.flops, "bytes": st.bytes}
                                         for st in plan.steps]}, open(out, "w"))


if __name__ == "__main__":
    main()

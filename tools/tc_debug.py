"""Per-role cycle counters of every tcgen05 conv launch of one 128^3 patch (VSSEG_TC_DEBUG=1):
   VSSEG_TC_DEBUG=1 python tools/tc_debug.py 2> tc_debug.log"""
import os
import sys

os.environ.setdefault("VSSEG_TC_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from vs_seg_b200 import lib as L
from vs_seg_b200 import sliding_window as sw
from vs_seg_b200.tensors import f32view


def main():
    dev = torch.device("cuda:0")
    net, _ = bench.build_net(dev)
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    plan = net.eval_plan(bench.ROI, batch, dev)
    vol = torch.randn((batch, 1) + bench.ROI, device=dev)
    out = torch.zeros((batch, 2) + bench.ROI, device=dev)
    plan._set(plan.src, f32view(vol))
    plan._set(plan.dst, f32view(out))
    s = torch.cuda.current_stream(dev).cuda_stream
    for it in range(2):
        for st in plan.steps:
            if it:
                sys.stderr.write(f"== {st.name}\n")
                sys.stderr.flush()
            L.check(st.fn(*st.args, s), st.name)
        torch.cuda.synchronize()
        if not it:
            sys.stderr.write("#### second pass ####\n")


if __name__ == "__main__":
    main()

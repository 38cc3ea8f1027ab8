"""Experiment (GPU box): 4 windows as ONE group of 4 vs TWO groups of 2 on two streams (atomic blend), the per-rank
work of the 8-GPU benchmark.  python tools/exp_split_groups.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from vs_seg_b200 import sliding_window as sw
from vs_seg_b200.tensors import f32view

dev = torch.device("cuda:0")
net, _ = bench.build_net(dev)
roi = bench.ROI
vol = torch.randn((1, 1) + bench.VOLUME, device=dev)
acc = torch.zeros((1, 2) + bench.VOLUME, device=dev)
imap = sw.importance_map(roi, "gaussian", 0.125, dev)
starts = sw.window_starts(bench.VOLUME, roi, 0.25)[:4]
src = [f32view(vol, s, roi) for s in starts]
dst = [f32view(acc, s, roi) for s in starts]
side = torch.cuda.Stream(dev)

def one_group():
    net.eval_plan(roi, 4, dev, window_levels=1).run(src, dst, imap.data_ptr(), atomic=True)

def two_groups(n=2):
    cur = torch.cuda.current_stream(dev)
    ev = torch.cuda.Event(); ev.record(cur); side.wait_event(ev)
    with torch.cuda.stream(side):
        net.eval_plan(roi, n, dev, window_levels=1, slot=1).run(src[n:], dst[n:], imap.data_ptr(), atomic=True)
        j = torch.cuda.Event(); j.record(side)
    net.eval_plan(roi, n, dev, window_levels=1, slot=0).run(src[:n], dst[:n], imap.data_ptr(), atomic=True)
    cur.wait_event(j)

def graphed(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        fn()
    return g

for name, fn in (("one group of 4", one_group), ("two groups of 2, two streams", two_groups)):
    g = graphed(fn)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20:.3f} ms per 4 windows")

#!/usr/bin/env python
"""bench.py — patches/sec (128^3) of sliding-window inference on synthetic 384x384x160 volumes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A step = the sliding-window inference of ONE synthetic volume (BASELINE.json configs[3]:
384x384x160, 128^3 window, overlap 0.25, gaussian blending -> 32 patches) including the finalise
kernel (probabilities, argmax mask, Dice sums).  N>1 (torchrun): the 32 windows of every volume are
sharded by patch index over the ranks, which blend them straight into rank 0's accumulator over NVLink
peer memory (strong scaling of a fixed volume).  One JSON line is printed by rank 0.

--impl reference times the CPU oracle (restatement of the reference's torch/MONAI path; the
reference itself cannot be imported on the GPU box) on the host cores, one patch per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOLUME = (384, 384, 160)
ROI = (128, 128, 128)
METRIC = "patches/sec (128^3) sliding-window infer"
N_ROT = 4  # rotating volume buffers: 4 x (94 MB in + 189 MB acc) >> 126 MB L2


def synth_volume(i, shape=VOLUME):
    """N(0,1) noise + bright ellipsoid 'tumour', whole-volume normalised; label = mask."""
    g = torch.Generator().manual_seed(1000 + i)
    x = torch.randn(shape, generator=g)
    c = [(0.25 + 0.5 * torch.rand(1, generator=g).item()) * s for s in shape]
    xs = torch.arange(shape[0]).view(-1, 1, 1).float()
    ys = torch.arange(shape[1]).view(1, -1, 1).float()
    zs = torch.arange(shape[2]).view(1, 1, -1).float()
    m = (((xs - c[0]) / 14) ** 2 + ((ys - c[1]) / 14) ** 2 + ((zs - c[2]) / 8) ** 2) <= 1
    x = x + 2.0 * m.float()
    x = (x - x.mean()) / x.std()
    return x[None, None].contiguous(), m.float()[None, None].contiguous()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [v for v in sm if v > 0.5 * (mx or 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_net(device):
    from oracle.unet_oracle import (CHANNELS, KERNEL_SIZES, SAMPLE_KERNEL_SIZES, STRIDES, seeded_state_dict)
    from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA
    sd = seeded_state_dict(0)  # random-init weights of the reference architecture (no checkpoints offline)
    net = UNet2d5_spvPA(dimensions=3, in_channels=1, out_channels=2, channels=CHANNELS, strides=STRIDES,
                        kernel_sizes=KERNEL_SIZES, sample_kernel_sizes=SAMPLE_KERNEL_SIZES, num_res_units=2,
                        norm="BATCH", dropout=0.1)
    net.load_state_dict(sd)
    return net.to(device).eval(), sd


def bench_config(world):
    """The workload description both arms print (the driver compares the two dicts)."""
    group = int(os.environ.get("VSSEG_SW_GROUP", "8"))
    return {"workload": "VS_inference sliding-window 384x384x160, 128^3 window (configs[3])",
            "volume": list(VOLUME), "roi": list(ROI), "overlap": 0.25, "blend": "gaussian",
            "patches_per_step": 32, "weights": "seeded random init",
            "l2_policy": f"GPU arm: {N_ROT} rotating volumes (283 MB each) > 126 MB L2",
            "gpu_schedule": f"window groups of {group}, one captured CUDA graph per volume; "
                            + (f"patch-index shard x{world}, blended into rank 0's accumulator over NVLink peer memory"
                               if world > 1 else "1 GPU")}


def train_record(dev, world, steps=5):
    """Secondary record (BASELINE configs[2], and configs[4] under torchrun): one training step = native train-mode
    forward + Dice_spvPA + native backward (+ ONE all-reduce of the flat gradient at N > 1) + fused Adam on a
    synthetic batch of 2 x 128^3 per GPU.  CUDA events, max over ranks."""
    import torch.distributed as dist
    from oracle.unet_oracle import CHANNELS, KERNEL_SIZES, SAMPLE_KERNEL_SIZES, STRIDES
    from params.losses.dice_spvPA import Dice_spvPA
    from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA
    from vs_seg_b200 import ddp
    from vs_seg_b200 import lib as vlib
    from vs_seg_b200.optim import FusedAdam
    torch.manual_seed(0)
    net = UNet2d5_spvPA(dimensions=3, in_channels=1, out_channels=2, channels=CHANNELS, strides=STRIDES,
                        kernel_sizes=KERNEL_SIZES, sample_kernel_sizes=SAMPLE_KERNEL_SIZES, num_res_units=2,
                        norm="BATCH", dropout=0.1).to(dev).train()
    ddp.broadcast_module_state(net)
    opt = FusedAdam(net.parameters(), lr=1e-4, weight_decay=1e-7)
    red = ddp.GradReducer(net, opt)
    crit = Dice_spvPA(to_onehot_y=True, softmax=True)
    g = torch.Generator().manual_seed(2000 + (dist.get_rank() if world > 1 else 0))
    x = torch.randn((2, 1) + ROI, generator=g).to(dev)
    y = (torch.rand((2, 1) + ROI, generator=g) > 0.95).float().to(dev)

    def step():
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        red.reduce()
        opt.step()
        return loss.detach()

    from vs_seg_b200.training import GraphedTrainStep
    graphed = GraphedTrainStep(net, crit, opt, red)

    def measure(fn, n):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        l0 = vlib.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            loss = fn()
        e1.record()
        host = (time.perf_counter() - t0) * 1e3 / n   # host time to ISSUE a step (no sync inside)
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), host, (vlib.launches() - l0) // n, loss

    for _ in range(2):
        step()
    eager_ms, eager_host, eager_launches, _ = measure(step, 3)
    graphed(x, y)   # capture (includes its own warm-up; the training state is restored around it)
    graphed(x, y)
    ms_v, host_ms, _, loss = measure(lambda: graphed(x, y), steps)
    ms = torch.tensor([ms_v])
    rec = {"workload": "training step, batch 2 x 128^3 per GPU: native fwd + Dice_spvPA + native bwd"
                       + (f" + 1 NCCL all-reduce of the flat gradient over {world} ranks" if world > 1 else "")
                       + " + fused Adam (BASELINE configs[2]" + ("/[4])" if world > 1 else ")")
                       + ", the whole step replayed as one CUDA graph (GraphedTrainStep, what VS_train.py runs)",
           "n_gpus": world, "batch_per_gpu": 2, "ms_per_step": ms.item(), "samples_per_s": 2 * world / (ms.item() * 1e-3),
           "host_issue_ms_per_step": host_ms,
           "eager": {"ms_per_step": eager_ms, "host_issue_ms_per_step": eager_host, "native_launches_per_step": eager_launches},
           "loss": float(loss), "dropout": 0.1, "steps": steps}
    del net, opt, red, graphed
    torch.cuda.empty_cache()
    return rec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.unet_oracle import seeded_state_dict
    sd = seeded_state_dict(0)
    vol, _ = synth_volume(0)
    from oracle import sw_oracle, unet_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    starts = sw_oracle.window_starts(VOLUME, ROI, sw_oracle.scan_interval(VOLUME, ROI, 0.25))
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            s = starts[i % len(starts)]
            w = vol[:, :, s[0]:s[0] + ROI[0], s[1]:s[1] + ROI[1], s[2]:s[2] + ROI[2]]
            t0 = time.perf_counter()
            unet_oracle.unet_forward(sd, w)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    val = len(times) / total
    sample = f"{len(times)} patches (128^3) of synthetic volume 0, 1 patch per step, sw_batch_size=1"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "patches/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step record")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from vs_seg_b200 import lib as vlib
    from vs_seg_b200 import parallel as par
    from vs_seg_b200 import sliding_window as sw
    from vs_seg_b200.tensors import f32view

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vlib.load()
    net, sd = build_net(dev)
    predictor = lambda t: net(t)[0]  # noqa: E731
    predictor.native_model = net

    host = [synth_volume(i) for i in range(N_ROT)]
    # the label travels as uint8 (1 byte per voxel; the finalise kernel reads it as bytes)
    pinned = [(v.pin_memory(), l.to(torch.uint8).pin_memory()) for v, l in host]
    vols = [(v.to(dev), l.to(torch.uint8).to(dev)) for v, l in host]
    n_win = len(sw.window_starts(VOLUME, ROI, 0.25))
    mask_host = torch.empty((1, 1) + VOLUME, dtype=torch.uint8).pin_memory()
    sums_host = torch.empty((1, 3), dtype=torch.float64).pin_memory()

    def infer(vol, label):
        # windows sharded by index over the ranks, blended into rank 0's accumulator over peer memory
        return par.sharded_sliding_window_inference(vol, ROI, 1, predictor, 0.25, "gaussian", label=label,
                                                    return_mask=True)

    def step_resident(i):
        v, l = vols[i % N_ROT]
        return infer(v, l)

    dv = [torch.empty_like(vols[0][0]) for _ in range(2)]
    dl = [torch.empty_like(vols[0][1]) for _ in range(2)]

    # a rank only needs the x slab its windows cover (windows are sharded in x-slowest order); the label is
    # only read by rank 0's finalise kernel
    slab = par.shard_slab(VOLUME, ROI, 0.25, rank, world) or (0, 0)
    h2d_bytes = (slab[1] - slab[0]) * VOLUME[1] * VOLUME[2] * 4 + (host[0][1].numel() if rank == 0 else 0)

    # end-to-end step: the volume comes from pinned host memory and the mask + Dice sums go back to the host
    # every step.  The copy of step i+1's input is issued on a copy stream while step i computes (two device
    # buffers); every step still pays one host->device and one device->host copy inside the timed region.
    copy_stream = torch.cuda.Stream(dev)
    d2h_stream = torch.cuda.Stream(dev)   # results leave on their own stream: the next volume's kernels do not wait
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    state = {"prefetched": None}

    def prefetch(j):
        hv, hl = pinned[j % N_ROT]
        v, l = dv[j % 2], dl[j % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[j % 2])   # the step that last read this buffer pair has finished
            v[:, :, slab[0]:slab[1]].copy_(hv[:, :, slab[0]:slab[1]], non_blocking=True)
            if rank == 0:
                l.copy_(hl, non_blocking=True)
            ready[j % 2].record(copy_stream)
        state["prefetched"] = j

    def step_e2e(i):
        cur = torch.cuda.current_stream(dev)
        if state["prefetched"] != i:
            prefetch(i)
        cur.wait_event(ready[i % 2])
        prefetch(i + 1)   # buffer (i+1) % 2 was released by step i-1
        res = infer(dv[i % 2], dl[i % 2])
        free[i % 2].record(cur)
        if res is not None:
            _, mask, sums = res
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                mask_host.copy_(mask, non_blocking=True)
                sums_host.copy_(sums, non_blocking=True)
            mask.record_stream(d2h_stream)
            sums.record_stream(d2h_stream)

    def timed(fn, warmup, steps, finish=None):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        l0 = vlib.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if finish is not None:
            finish()   # the timed region ends when the last result has reached the host
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), vlib.launches() - l0

    with torch.no_grad():
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms, launches = timed(step_resident, max(args.warmup, 3), args.steps)
        ms_e2e, _ = timed(step_e2e, 2, args.steps, finish=lambda: torch.cuda.current_stream(dev).wait_stream(d2h_stream))
        clk = clocks.stop() if rank == 0 else None

        # parity leg, part 1 (all ranks): the SAME call the timed region makes, on synthetic volume 0 - window
        # groups of 8 through the captured graph, blend in the last kernel, (N > 1: NCCL reduce), finalise
        par_res = None
        if not args.no_cpu_baseline:
            res = infer(vols[0][0], vols[0][1])
            if res is not None:
                par_res = (res[0].cpu(), res[1].cpu(), res[2].cpu())
            if world > 1 and par._PEER:
                next(iter(par._PEER.values())).check()   # no rank timed out in the peer hand-shake
            torch.cuda.synchronize(dev)

    train = None
    if not args.no_train:
        try:
            train = train_record(dev, world)
        except Exception as e:  # noqa: BLE001  (secondary record: never lose the headline line)
            train = {"error": repr(e)[:300]}

    with torch.no_grad():
        if rank != 0:
            dist.destroy_process_group()
            return
        value = n_win * args.steps / (ms * 1e-3)
        e2e_value = n_win * args.steps / (ms_e2e * 1e-3)

        # ---- per-launch profile of one window group (outside the timed region) -> roofline of the top launch.
        # Same plan as the timed region uses.  A launch is tensor-bound when 3 x FLOP / tensor peak exceeds
        # bytes / HBM peak (every product is three bf16 MMAs: hi*hi + lo*hi + hi*lo).
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        group = min(int(os.environ.get("VSSEG_SW_GROUP", "8")), n_win)
        levels = int(os.environ.get("VSSEG_SW_WINDOW_LEVELS", "1"))
        plan = net.eval_plan(ROI, batch=group, device=dev, window_levels=levels)
        vol0 = vols[0][0]
        acc = torch.zeros((1, 2) + VOLUME, device=dev)
        imap = sw.importance_map(ROI, "gaussian", 0.125, dev)
        starts = sw.window_starts(VOLUME, ROI, 0.25)[:group]
        prof = plan.profile([f32view(vol0, s, ROI) for s in starts], [f32view(acc, s, ROI) for s in starts],
                            imap.data_ptr())
        group_ms = sum(p[4] for p in prof)
        patch_ms = group_ms / group
        pk_t = peaks.get("bf16_tflops", 1590.0) * 1e12   # a launch timed alone: the burst figure
        pk_h = peaks.get("hbm_gbs", 6650.0) * 1e9
        top = max(prof, key=lambda p: p[4])
        t_tensor, t_hbm = 3 * top[2] / pk_t, top[3] / pk_h
        if t_tensor > t_hbm:
            roof = {"bound": "tensor", "achieved": 3 * top[2] / (top[4] * 1e-3) / 1e12, "peak": pk_t / 1e12,
                    "unit": "TFLOP/s", "note": "issued bf16 MMA FLOPs = 3 x algorithmic (bf16x3); useful = achieved / 3"}
        else:
            roof = {"bound": "hbm", "achieved": top[3] / (top[4] * 1e-3) / 1e9, "peak": pk_h / 1e9, "unit": "GB/s"}
        traffic = None
        try:   # dram bytes of the same launch from the committed ncu --set full capture (profiles/)
            import glob
            ncu = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_group_summary.json")))[-1]))
            hit = [e for e in ncu["launches"] if e["step"] == top[0]]
            if hit and hit[0].get("dram_rd_MB") is not None:
                traffic = (hit[0]["dram_rd_MB"] + hit[0]["dram_wr_MB"]) * 1e6
        except (OSError, KeyError, ValueError):
            pass
        tc_ms = sum(p[4] for p in prof if p[1] == "tcgen05")
        # per-launch roofline time: max(3 x FLOP / tensor peak, algorithmic bytes / HBM peak)  (bf16x3 issues three
        # MMAs per product; bytes follow SURVEY.md §8d: every activation read once per consumer and written once, 4 B)
        roof_ms = sum(max(3 * p[2] / pk_t, p[3] / pk_h) for p in prof) * 1e3
        timed_patch_ms = ms / args.steps / (n_win / world)
        roof.update({"frac": roof["achieved"] / roof["peak"], "traffic": traffic, "kernel": "conv_tc_kernel" if top[1] == "tcgen05" else top[0],
                     "launch": top[0], "alg_bytes_per_launch": top[3], "alg_flops_per_launch": top[2],
                     "kernel_ms": top[4], "share_of_group": top[4] / group_ms,
                     "conv_tc_kernel_share_of_group": tc_ms / group_ms,
                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                     "per_launch_frac_definition": "max(3*FLOP/tensor_peak, alg_bytes/hbm_peak) / kernel_ms",
                     "whole_patch": {
                         "alg_gflop": plan.total_flops() / group / 1e9,
                         "alg_gbytes": sum(p[3] for p in prof) / group / 1e9,
                         "roofline_ms": roof_ms / group,
                         "survey_roofline_ms_bf16x3_4B": 0.755,
                         "profiled_ms": patch_ms, "timed_ms": timed_patch_ms,
                         "frac_profiled": roof_ms / group_ms, "frac_timed": roof_ms / group / timed_patch_ms,
                         "frac_timed_vs_survey": 0.755 / timed_patch_ms,
                         "useful_tflops": plan.total_flops() / (group_ms * 1e-3) / 1e12},
                     "launches": [{"launch": p[0], "ms": round(p[4], 4), "gflop": round(p[2] / 1e9, 3),
                                   "mbytes": round(p[3] / 1e6, 1),
                                   "frac": round(max(3 * p[2] / pk_t, p[3] / pk_h) * 1e3 / p[4], 3)}
                                  for p in sorted(prof, key=lambda q: -q[4])[:12]]})

        # ---- CPU baseline (the oracle on the host cores) + parity of the timed configuration against it:
        # the whole finalised 384x384x160 volume of the call above vs sw_oracle driving the oracle network over
        # all 32 windows (sw_batch_size 1, as VSparams.py:568-574); the oracle's per-window time is the CPU baseline
        cpu = None
        parity = None
        if not args.no_cpu_baseline and par_res is not None:
            from oracle import loss_oracle, sw_oracle, unet_oracle
            torch.set_num_threads(os.cpu_count() or 1)
            times = []

            def oracle_predictor(w):
                t0 = time.perf_counter()
                y = unet_oracle.unet_forward(sd, w)[0]
                times.append(time.perf_counter() - t0)
                return y

            with torch.no_grad():
                ref = sw_oracle.sliding_window_inference(host[0][0], ROI, 1, oracle_predictor, mode="gaussian")
            timed_t = times[1:] if len(times) > 1 else times   # the first window is the warm-up
            cpu = {"value": len(timed_t) / sum(timed_t), "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{len(timed_t)} consecutive 128^3 windows of synthetic volume 0 after 1 warm-up window "
                             f"({sum(timed_t):.1f} s), torch {torch.__version__} fp32, {os.cpu_count()} host cpus"}
            got, mask, sums = par_res
            label = host[0][1]
            margin = (ref[:, 1] - ref[:, 0]).abs()
            ref_mask = ref.argmax(1, keepdim=True)
            diff_mask = mask.long() != ref_mask
            dice_native = ((2 * sums[0, 0] + 1e-5) / (sums[0, 1] + sums[0, 2] + 1e-5)).item()
            parity = {"scope": "full finalised volume of the timed call (window groups of 8, captured graph, blend in the "
                               "last kernel" + (f", peer-memory blend of {world} ranks" if world > 1 else "") + ") vs sw_oracle",
                      "max_abs_err_logits": (got - ref).abs().max().item(),
                      "argmax_flips_margin_gt_1e-4": ((got.argmax(1) != ref.argmax(1)) & (margin > 1e-4)).sum().item(),
                      "mask_kernel_flips_margin_gt_1e-4": (diff_mask[:, 0] & (margin > 1e-4)).sum().item(),
                      "near_ties_le_1e-4": (margin <= 1e-4).sum().item(),
                      "dice_mask_native": dice_native,
                      "dice_mask_oracle": loss_oracle.dice_score(ref, label).item(),
                      "windows_checked": len(times), "voxels_checked": ref[:, 0].numel(), "tolerance": 1e-3}

        in_bytes = h2d_bytes
        out_bytes = mask_host.numel() + sums_host.numel() * 8
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16x3->f32",
            "data": "synthetic",
            "config": bench_config(world),
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "patch_ms_profiled": patch_ms, "train": train,
            "notes": "value/e2e: CUDA events around K volumes, max over ranks; e2e copies the rank's x slab of the "
                     "volume (and the label on rank 0) from pinned host memory and the mask + Dice sums back every step; "
                     "h2d_bytes_per_step is rank 0's",
        }))
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

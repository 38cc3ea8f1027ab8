"""BASELINE configs[1]: the widest encoder ResidualUnit (level 3: Conv3d 32->48 + 48->48, k(3,3,3), BatchNorm(eval) +
PReLU, 1x1x1 shortcut, on [2,32,32,32,128] - SURVEY.md §8 rows 8-10) - native fused launches vs torch/cuDNN eager.
    python tools/bench_block.py            (GPU box)
Native numbers are the two launches enc2.unit0 / enc2.unit1 of a batch-2 eval plan (the shortcut is a second
accumulator of unit1), timed per launch with CUDA events; cuDNN runs the same torch modules in eval mode."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vs_seg_b200.tensors import f32view  # noqa: E402


def cudnn_block(sd, dev):
    import torch.nn as nn
    p = "model.1.submodule.1.1.submodule.1.0."

    class RU(nn.Module):
        def __init__(self):
            super().__init__()
            self.c0, self.b0, self.a0 = nn.Conv3d(32, 48, 3, padding=1), nn.BatchNorm3d(48), nn.PReLU()
            self.c1, self.b1, self.a1 = nn.Conv3d(48, 48, 3, padding=1), nn.BatchNorm3d(48), nn.PReLU()
            self.res = nn.Conv3d(32, 48, 1)

        def forward(self, x):
            h = self.a0(self.b0(self.c0(x)))
            return self.a1(self.b1(self.c1(h))) + self.res(x)

    m = RU()
    with torch.no_grad():
        for i, (c, b, a) in enumerate(((m.c0, m.b0, m.a0), (m.c1, m.b1, m.a1))):
            q = p + f"conv.unit{i}."
            c.weight.copy_(sd[q + "conv.weight"]); c.bias.copy_(sd[q + "conv.bias"])
            b.weight.copy_(sd[q + "norm.weight"]); b.bias.copy_(sd[q + "norm.bias"])
            b.running_mean.copy_(sd[q + "norm.running_mean"]); b.running_var.copy_(sd[q + "norm.running_var"])
            a.weight.copy_(sd[q + "act.weight"])
        m.res.weight.copy_(sd[p + "residual.weight"]); m.res.bias.copy_(sd[p + "residual.bias"])
    return m.to(dev).eval()


def time_fn(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    net, sd = bench.build_net(dev)
    B = 2
    plan = net.eval_plan(bench.ROI, B, dev)
    x = torch.randn((B, 1) + bench.ROI, device=dev)
    out = torch.empty((B, 2) + bench.ROI, device=dev)
    prof = {p[0]: p for p in plan.profile(f32view(x), f32view(out), None, iters=10)}
    u0, u1 = prof["enc2.unit0"], prof["enc2.unit1"]
    flops = u0[2] + u1[2]
    ms = u0[4] + u1[4]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    pk = peaks.get("bf16_tflops", 1590.0)
    res = {"block": "encoder level-3 ResidualUnit 32->48->48 k(3,3,3) + 1x1x1 shortcut, input [2,32,32,32,128], eval",
           "gflop": flops / 1e9, "native_ms": ms, "native_launches": 2,
           "native_useful_tflops": flops / ms / 1e9, "native_issued_bf16_tflops": 3 * flops / ms / 1e9,
           "issued_frac_of_bf16_peak": 3 * flops / ms / 1e9 / pk, "bf16_peak_tflops": pk}
    blk = cudnn_block({k: v.to(dev) for k, v in sd.items()}, dev)
    xin = torch.randn((B, 32, 32, 32, 128), device=dev)
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        res["cudnn_fp32_ms"] = time_fn(lambda: blk(xin))
        torch.backends.cudnn.allow_tf32 = True
        res["cudnn_tf32_ms"] = time_fn(lambda: blk(xin))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["cudnn_bf16_autocast_ms"] = time_fn(lambda: blk(xin))
        blk_cl = blk.to(memory_format=torch.channels_last_3d)
        xcl = xin.to(memory_format=torch.channels_last_3d)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["cudnn_bf16_channels_last_ms"] = time_fn(lambda: blk_cl(xcl))
    res["train"] = train_leg(net, dev, blk)
    print(json.dumps(res))


def train_leg(net, dev, blk):
    """Backward leg of configs[1]: the same ResidualUnit in TRAIN mode (batch-stat BatchNorm, dropout 0.1, PReLU) -
    native forward tape + backward (tcgen05 data gradients and weight gradients, BN/PReLU backward) vs torch/cuDNN
    autograd.  The native numbers are CUDA-event times around UNetTrainStep.residual_unit and its backward closure."""
    from vs_seg_b200.tensors import Act8Buffer
    from vs_seg_b200.training import UNetTrainStep, _GradBuf
    B, dims = 2, (32, 32, 128)
    prefix = "model.1.submodule.1.1.submodule.1.0."
    net.train()
    x1 = torch.randn((B, 1) + bench.ROI, device=dev)
    xin = torch.randn((B, 32) + dims, device=dev)
    dout_t = torch.randn((B, 48) + dims, device=dev)
    out = {}

    def once():
        st = UNetTrainStep(net, x1)
        src = Act8Buffer(B, 32, *dims, dev).from_ncdhw(xin)
        gsrc = _GradBuf(src)
        dst = Act8Buffer(B, 48, *dims, dev)
        dout = Act8Buffer(B, 48, *dims, dev).from_ncdhw(dout_t)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        bw = st.residual_unit(prefix, src.view(), gsrc, dst, 0, 48, (3, 3, 3), 2)
        e[1].record()
        bw(dout.view())
        e[2].record()
        torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])

    for _ in range(2):
        once()
    runs = [once() for _ in range(5)]
    out["native_fwd_ms"] = sorted(r[0] for r in runs)[len(runs) // 2]
    out["native_bwd_ms"] = sorted(r[1] for r in runs)[len(runs) // 2]
    out["note"] = ("native times include the host issuing the ~25 launches of the tape (CUDA events on the stream); "
                   "fwd = conv x3 + BN stats/finalise/apply x2, bwd = BN/PReLU backward + wgrad + dgrad per conv")
    net.eval()
    blk = blk.float().train()
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.momentum = 0.1
    xg = xin.clone().requires_grad_(True)

    def eager():
        y = blk(xg)
        y.backward(dout_t)

    for name, tf32, cast in (("cudnn_fp32", False, None), ("cudnn_tf32", True, None), ("cudnn_bf16_autocast", True, torch.bfloat16)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            if cast is None:
                out[name + "_fwd_bwd_ms"] = time_fn(eager, n=10)
            else:
                def eager_cast():
                    with torch.autocast("cuda", dtype=cast):
                        y = blk(xg)
                    y.backward(dout_t.to(y.dtype))
                out[name + "_fwd_bwd_ms"] = time_fn(eager_cast, n=10)
        except Exception as e:  # noqa: BLE001
            out[name + "_error"] = repr(e)[:200]
    out["native_fwd_bwd_ms"] = out["native_fwd_ms"] + out["native_bwd_ms"]
    return out


if __name__ == "__main__":
    main()

"""Training-step timing (BASELINE configs[2]: full attention-UNet step, batch 2, 128^3 crop, 1xB200):
native path vs the torch/cuDNN eager composition of the same modules (the incumbent on this box).
usage (GPU box): python tools/bench_train.py [B X Y Z] [--steps N]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle  # noqa: E402  (seeded weights only)
from params.losses.dice_spvPA import Dice_spvPA  # noqa: E402
from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA  # noqa: E402
from vs_seg_b200 import lib as vlib  # noqa: E402


def make(dev):
    torch.manual_seed(0)
    net = UNet2d5_spvPA(dimensions=3, in_channels=1, out_channels=2, channels=unet_oracle.CHANNELS,
                        strides=unet_oracle.STRIDES, kernel_sizes=unet_oracle.KERNEL_SIZES,
                        sample_kernel_sizes=unet_oracle.SAMPLE_KERNEL_SIZES, num_res_units=2, norm="BATCH", dropout=0.1)
    return net.to(dev).train()


def timed(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shape = tuple(int(v) for v in args[:4]) if len(args) >= 4 else (2, 128, 128, 128)
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 3
    dev = torch.device("cuda:0")
    B = shape[0]
    g = torch.Generator().manual_seed(1)
    x = torch.randn((B, 1) + shape[1:], generator=g).to(dev)
    y = (torch.rand((B, 1) + shape[1:], generator=g) > 0.95).float().to(dev)
    crit = Dice_spvPA(to_onehot_y=True, softmax=True)
    out = {"shape": list(shape), "steps": steps}

    net = make(dev)
    from vs_seg_b200.optim import FusedAdam
    opt = FusedAdam(net.parameters(), lr=1e-4, weight_decay=1e-7)   # one fused launch over the flat parameter buffer
    if os.environ.get("VSSEG_BENCH_TORCH_ADAM"):
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=1e-7)

    def native_step():
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        return loss

    l0 = vlib.launches()
    ms, wall = timed(native_step, 1, steps)
    out["native_ms_per_step"], out["native_wall_ms"] = ms, wall
    out["native_launches_per_step"] = (vlib.launches() - l0) // (steps + 1)
    out["native_loss"] = native_step().item()
    out["native_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    del net, opt
    torch.cuda.empty_cache()
    if "--native-only" in sys.argv:
        print(json.dumps(out))
        return

    # incumbent: the same module tree executed by torch/cuDNN eager.  The drop-in modules dispatch CUDA tensors to
    # the native kernels, so this TOOL (not the product) re-points their forward methods at the plain torch containers.
    from params.networks.blocks import attentionblock as ab
    from params.networks.blocks import convolutions as cv
    cv.Convolution.forward = lambda self, t: torch.nn.Sequential.forward(self, t)
    cv.ResidualUnit.forward = lambda self, t: self.conv(t) + self.residual(t)
    ab.AttentionBlock2.forward = lambda self, tup: tup[0].repeat([1, self.in_channels, 1, 1, 1]) * tup[1] + tup[1]
    ref = make(dev)
    opt2 = torch.optim.Adam(ref.parameters(), lr=1e-4, weight_decay=1e-7)

    def eager_step():
        opt2.zero_grad()
        ref.att_maps = []
        logits = ref.model(x)            # bypasses the native dispatch: plain torch modules on cuDNN
        loss = crit((logits, ref.att_maps), y)
        loss.backward()
        opt2.step()
        return loss

    try:
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms2, wall2 = timed(eager_step, 1, steps)
            out["cudnn_eager_ms_per_step_tf32" if tf32 else "cudnn_eager_ms_per_step_fp32"] = ms2
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms2, wall2 = timed(eager_step, 1, steps)
            out["cudnn_eager_ms_per_step_bf16_autocast"] = ms2
    except Exception as e:  # noqa: BLE001
        out["cudnn_eager_error"] = repr(e)[:200]
    print(json.dumps(out))


if __name__ == "__main__":
    main()

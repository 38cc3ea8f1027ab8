"""Data-parallel training check / timing on N GPUs (BASELINE configs[4] shape: batch 2 per GPU, 128^3 crop):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_train_check.py [steps]
Every rank: native forward + Dice_spvPA + native backward, ONE all-reduce of the flat gradient, fused Adam.
Prints step time (max over ranks, CUDA events) and verifies that the replicas' parameters stay identical."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from params.losses.dice_spvPA import Dice_spvPA  # noqa: E402
from vs_seg_b200 import ddp  # noqa: E402
from vs_seg_b200.optim import FusedAdam  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    crop = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (128, 128, 128)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)   # per-rank init: the broadcast must make the replicas equal
    net, _ = bench.build_net(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.01 * rank)
    net.train()
    ddp.broadcast_module_state(net)
    opt = FusedAdam(net.parameters(), lr=1e-4, weight_decay=1e-7)
    red = ddp.GradReducer(net, opt)
    crit = Dice_spvPA(to_onehot_y=True, softmax=True)
    g = torch.Generator().manual_seed(2000 + rank)     # every rank its own shard of synthetic crops
    x = torch.randn((2, 1) + crop, generator=g).to(dev)
    y = (torch.rand((2, 1) + crop, generator=g) > 0.95).float().to(dev)

    def step():
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        red.reduce()
        opt.step()
        return loss

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([float(torch.equal(ref, flat))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    if rank == 0:
        print(json.dumps({"what": "data-parallel training step (fwd + Dice_spvPA + bwd + grad all-reduce + fused Adam)",
                          "n_gpus": world, "batch_per_gpu": 2, "crop": list(crop), "ms_per_step": ms.item(),
                          "samples_per_s": 2 * world / (ms.item() * 1e-3), "replicas_identical": same,
                          "loss": float(loss)}))
    assert same, "replicas diverged"
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

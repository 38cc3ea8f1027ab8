"""Worker of tests/test_gpu_round2.py::test_two_rank_nccl_sharded_inference_equals_one_gpu (launched by torchrun,
one rank per GPU): sharded sliding-window inference over NCCL vs the same volume on one GPU and vs the oracle."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import sw_oracle, unet_oracle  # noqa: E402


def main():
    out_path = sys.argv[1]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA
    from vs_seg_b200 import parallel as par
    from vs_seg_b200 import sliding_window as sw
    sd = unet_oracle.seeded_state_dict(4)
    net = UNet2d5_spvPA(dimensions=3, in_channels=1, out_channels=2, channels=unet_oracle.CHANNELS,
                        strides=unet_oracle.STRIDES, kernel_sizes=unet_oracle.KERNEL_SIZES,
                        sample_kernel_sizes=unet_oracle.SAMPLE_KERNEL_SIZES, num_res_units=2, norm="BATCH", dropout=0.1)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    roi = (64, 64, 16)
    g = torch.Generator().manual_seed(33)
    x = torch.randn((1, 1, 160, 96, 24), generator=g)
    label = (torch.rand((1, 1, 160, 96, 24), generator=g) > 0.5).to(torch.uint8)
    with torch.no_grad():
        res = par.sharded_sliding_window_inference(x.to(dev), roi, 1, net, mode="gaussian", label=label.to(dev),
                                                   return_mask=True)
        # more calls through the cached program (captured graph): both buffers of the peer accumulator and the
        # release hand-shake (volume k waits for the finalise of volume k-2) are exercised
        for _ in range(4):
            res2 = par.sharded_sliding_window_inference(x.to(dev), roi, 1, net, mode="gaussian", label=label.to(dev),
                                                        return_mask=True)
    if dist.get_rank() == 0:
        out, mask, sums = res
        assert (out - res2[0]).abs().max().item() < 1e-5   # atomic blend over peer memory: sum order varies
        with torch.no_grad():
            acc, cnt, lows, img = sw.sliding_window_accumulate(x.to(dev), roi, net, mode="gaussian")
            single, mask1, sums1 = sw.finalize(acc, cnt, lows, img, label=label.to(dev), return_mask=True)
            ref = sw_oracle.sliding_window_inference(x, roi, 1, lambda w: unet_oracle.unet_forward(sd, w)[0], mode="gaussian")
        margin = (ref[:, 1] - ref[:, 0]).abs()
        torch.save({"err_vs_single": (out - single).abs().max().item(),
                    "err_vs_oracle": (out.cpu() - ref).abs().max().item(),
                    "flips": ((out.cpu().argmax(1) != ref.argmax(1)) & (margin > 1e-4)).sum().item(),
                    "mask_equal": bool(((mask != mask1).cpu()[:, 0] & (margin > 1e-4)).sum().item() == 0),
                    "dice_sharded": sw.dice_from_sums(sums)[0].item(), "dice_single": sw.dice_from_sums(sums1)[0].item()},
                   out_path)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

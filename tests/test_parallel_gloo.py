"""N>1 path on CPU: two gloo ranks shard the sliding-window windows and reduce the accumulator;
the result must equal the single-process result (and the oracle)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sw_oracle


def _predictor(x):
    # a cheap stand-in network with 2 output channels and a spatial footprint (so blending matters)
    k = torch.ones(2, 1, 3, 3, 3) / 27.0
    k[1] *= -0.5
    return torch.nn.functional.conv3d(x, k, padding=1) + x


_predictor.out_channels = 2


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vs_seg_b200.parallel import sharded_sliding_window_inference
    x = torch.randn((1, 1, 40, 36, 20), generator=torch.Generator().manual_seed(3))
    res = sharded_sliding_window_inference(x, (16, 16, 8), 1, _predictor, mode="gaussian")
    if rank == 0:
        torch.save(res, out)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_matches_single_process(tmp_path):
    from vs_seg_b200.sliding_window import shard_range, sliding_window_inference
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    x = torch.randn((1, 1, 40, 36, 20), generator=torch.Generator().manual_seed(3))
    single = sliding_window_inference(x, (16, 16, 8), 1, _predictor, mode="gaussian")
    ref = sw_oracle.sliding_window_inference(x, (16, 16, 8), 1, _predictor, mode="gaussian")
    assert got.shape == ref.shape
    assert (got - single).abs().max().item() < 1e-5
    assert (got - ref).abs().max().item() < 1e-5
    # the shards tile the window list exactly once
    for n, w in [(32, 8), (32, 3), (5, 8), (1, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_shard_slab_covers_the_shard_windows():
    from vs_seg_b200.parallel import shard_slab
    from vs_seg_b200.sliding_window import shard_range, window_starts
    img, roi = (384, 384, 160), (128, 128, 128)
    starts = window_starts(img, roi, 0.25)
    assert len(starts) == 32
    for world in (1, 2, 4, 8, 5):
        for r in range(world):
            lo, hi = shard_range(len(starts), r, world)
            x0, x1 = shard_slab(img, roi, 0.25, r, world)
            assert all(x0 <= s[0] and s[0] + roi[0] <= x1 for s in starts[lo:hi])
    assert shard_slab(img, roi, 0.25, 0, 8) == (0, 128) and shard_slab(img, roi, 0.25, 7, 8) == (256, 384)
    assert shard_slab((64, 64, 16), (64, 64, 16), 0.25, 0, 2) is None   # one window, two ranks: rank 1 owns it


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vs_seg_b200 import ddp
    from vs_seg_b200.optim import FusedAdam
    torch.manual_seed(100 + rank)           # different initial weights per rank: the broadcast must fix that
    net = torch.nn.Sequential(torch.nn.Conv3d(1, 4, 3, padding=1), torch.nn.BatchNorm3d(4), torch.nn.PReLU(),
                              torch.nn.Conv3d(4, 2, 1))
    ddp.broadcast_module_state(net)
    opt = FusedAdam(net.parameters(), lr=1e-2, weight_decay=1e-3)   # CPU parameters -> torch.optim.Adam inside
    red = ddp.GradReducer(net, opt)
    g = torch.Generator().manual_seed(7)
    xs = torch.randn((4, 1, 8, 8, 4), generator=g)                  # global batch of 4, 2 per rank
    x = xs[rank * 2:(rank + 1) * 2]
    for _ in range(3):
        opt.zero_grad()
        net(x).square().mean().backward()
        red.reduce()
        opt.step()
    torch.save({k: v.clone() for k, v in net.state_dict().items()}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_training_keeps_replicas_identical(tmp_path):
    """Two gloo ranks: broadcast of rank 0's weights, per-rank batches, one gradient all-reduce per step.
    The replicas' parameters must stay bit-identical (BatchNorm buffers are per-rank by design)."""
    from vs_seg_b200.ddp import shard_list
    out = str(tmp_path / "sd")
    mp.spawn(_ddp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = torch.load(out + ".0"), torch.load(out + ".1")
    for k in a:
        if "running_" in k or "num_batches" in k:
            continue
        assert torch.equal(a[k], b[k]), k
    # shards: padded round-robin, every item covered, equal lengths
    items = list(range(7))
    shards = [shard_list(items, r, 3) for r in range(3)]
    assert all(len(s) == 3 for s in shards) and set(sum(shards, [])) == set(items)
    assert shard_list(items, 0, 1) == items

"""Generate tests/golden/*.npz from the UNMODIFIED reference modules (build container only).

Run:  python oracle/make_golden.py      (needs /root/reference; not available on the GPU box)

The reference's network and loss files are imported as they lie under /root/reference with
oracle/monai_shim on sys.path (MONAI itself is not installable offline).  The seeded weights
come from oracle.unet_oracle.seeded_state_dict and are loaded with strict load_state_dict, which
also proves the key set / shapes equal the reference's.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "monai_shim"))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from oracle.unet_oracle import (CHANNELS, KERNEL_SIZES, SAMPLE_KERNEL_SIZES, STRIDES,  # noqa: E402
                                seeded_state_dict)


def load_reference():
    # `params` in sys.modules must be the reference's package, not this repo's mirror.
    for m in [m for m in sys.modules if m == "params" or m.startswith("params.")]:
        del sys.modules[m]
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "params", os.path.join(REF, "params", "__init__.py"),
        submodule_search_locations=[os.path.join(REF, "params")])
    if spec is None or not os.path.exists(os.path.join(REF, "params", "__init__.py")):
        import types
        pkg = types.ModuleType("params")
        pkg.__path__ = [os.path.join(REF, "params")]
        sys.modules["params"] = pkg
        for sub in ("networks", "networks.nets", "networks.blocks", "losses"):
            m = types.ModuleType("params." + sub)
            m.__path__ = [os.path.join(REF, "params", *sub.split("."))]
            sys.modules["params." + sub] = m
    from params.networks.nets.unet2d5_spvPA import UNet2d5_spvPA
    from params.losses.dice_spvPA import Dice_spvPA
    return UNet2d5_spvPA, Dice_spvPA


def build(UNet, attention, dropout=0.1):
    return UNet(dimensions=3, in_channels=1, out_channels=2, channels=CHANNELS, strides=STRIDES,
                kernel_sizes=KERNEL_SIZES, sample_kernel_sizes=SAMPLE_KERNEL_SIZES,
                num_res_units=2, norm="BATCH", dropout=dropout, attention_module=attention)


def synth_label(shape, g):
    """Ellipsoid 'tumour' label [B,1,X,Y,Z] (SURVEY.md §8d generator, scaled to the volume)."""
    b, _, X, Y, Z = shape
    xs = torch.arange(X).view(X, 1, 1).float()
    ys = torch.arange(Y).view(1, Y, 1).float()
    zs = torch.arange(Z).view(1, 1, Z).float()
    out = torch.zeros(shape)
    for i in range(b):
        c = [(0.25 + 0.5 * torch.rand(1, generator=g).item()) * s for s in (X, Y, Z)]
        r = (max(X / 6, 2), max(Y / 6, 2), max(Z / 5, 1.5))
        out[i, 0] = (((xs - c[0]) / r[0]) ** 2 + ((ys - c[1]) / r[1]) ** 2 + ((zs - c[2]) / r[2]) ** 2 <= 1).float()
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    UNet, DiceSpv = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # (i) eval forward, attention and no-attention
    for attention, shape, name in ((True, (1, 1, 64, 64, 16), "unet_eval_att"),
                                   (False, (1, 1, 32, 32, 8), "unet_eval_noatt")):
        sd = seeded_state_dict(seed=0, attention=attention)
        net = build(UNet, attention)
        assert list(net.state_dict().keys()) == list(sd.keys()), "key order/set differs"
        net.load_state_dict(sd, strict=True)
        net.eval()
        g = torch.Generator().manual_seed(123)
        x = torch.randn(shape, generator=g)
        with torch.no_grad():
            logits, atts = net(x)
        arrs = {"x": x.numpy(), "logits": logits.numpy(), "n_keys": np.array(len(sd)),
                "n_params": np.array(sum(p.numel() for p in net.parameters()))}
        for i, a in enumerate(atts):
            arrs[f"att{i}"] = a.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
        print(name, "logits", tuple(logits.shape), "keys", len(sd), "params", int(arrs["n_params"]))

    # (ii) train-mode forward + loss + backward (dropout = 0 so it is deterministic)
    sd = seeded_state_dict(seed=1, attention=True)
    net = build(UNet, True, dropout=0)
    net.load_state_dict(sd, strict=True)
    net.train()
    g = torch.Generator().manual_seed(321)
    x = torch.randn((2, 1, 32, 32, 16), generator=g)
    y = synth_label((2, 1, 32, 32, 16), g)
    loss_fn = DiceSpv(to_onehot_y=True, softmax=True, supervised_attention=True, hardness_weighting=True)
    logits, atts = net(x)
    loss = loss_fn((logits, atts), y)
    loss.backward()
    arrs = {"x": x.numpy(), "y": y.numpy(), "loss": np.array(loss.item()),
            "logits": logits.detach().numpy()}
    names, norms = [], []
    for n, p in net.named_parameters():
        names.append(n)
        norms.append(p.grad.double().norm().item())
    arrs["grad_names"] = np.array(names)
    arrs["grad_norms"] = np.array(norms)
    for n in ("model.0.residual.weight", "model.2.1.conv.unit0.conv.weight",
              "model.0.conv.unit0.norm.weight", "model.2.0.0.conv2.conv.weight"):
        arrs["grad::" + n] = dict(net.named_parameters())[n].grad.numpy()
    new_sd = net.state_dict()
    arrs["rm::model.0.conv.unit0.norm.running_mean"] = new_sd["model.0.conv.unit0.norm.running_mean"].numpy()
    arrs["rv::model.0.conv.unit0.norm.running_var"] = new_sd["model.0.conv.unit0.norm.running_var"].numpy()
    np.savez_compressed(os.path.join(out_dir, "unet_train_step.npz"), **arrs)
    print("train step loss", loss.item())

    # (iii) loss golden: seeded logits / att maps / labels incl. empty and full labels
    g = torch.Generator().manual_seed(7)
    shape = (2, 1, 32, 32, 16)
    att_shapes = [(2, 1, 1, 1, 2), (2, 1, 2, 2, 4), (2, 1, 4, 4, 8), (2, 1, 8, 8, 16),
                  (2, 1, 16, 16, 16), (2, 1, 32, 32, 16)]  # coarsest first
    cases = {}
    for cname in ("ellipsoid", "empty", "full"):
        x = (2.0 * torch.randn((2, 2, 32, 32, 16), generator=g)).requires_grad_(True)
        atts = [torch.rand(s, generator=g).requires_grad_(True) for s in att_shapes]
        if cname == "ellipsoid":
            y = synth_label(shape, g)
        elif cname == "empty":
            y = torch.zeros(shape)
        else:
            y = torch.ones(shape)
        for flags in ((True, True), (True, False), (False, True), (False, False)):
            fn = DiceSpv(to_onehot_y=True, softmax=True, supervised_attention=flags[0],
                         hardness_weighting=flags[1])
            for t in [x] + atts:
                t.grad = None
            loss = fn((x, atts), y)
            loss.backward()
            tag = f"{cname}_a{int(flags[0])}h{int(flags[1])}"
            cases[tag + "_loss"] = np.array(loss.item())
            if flags == (True, True):
                cases[cname + "_x"] = x.detach().numpy()
                cases[cname + "_y"] = y.numpy()
                cases[cname + "_gx"] = x.grad.numpy()
                for i, a in enumerate(atts):
                    cases[f"{cname}_att{i}"] = a.detach().numpy()
                    cases[f"{cname}_gatt{i}"] = a.grad.numpy()
    np.savez_compressed(os.path.join(out_dir, "dice_spvpa_loss.npz"), **cases)
    print("loss cases", sorted(k for k in cases if k.endswith("_loss")))


if __name__ == "__main__":
    main()

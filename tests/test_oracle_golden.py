"""The oracle restatement vs. golden vectors produced by the UNMODIFIED reference modules
(oracle/make_golden.py, run in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle, unet_oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("attention,name", [(True, "unet_eval_att.npz"), (False, "unet_eval_noatt.npz")])
def test_unet_eval_matches_reference(golden_dir, attention, name):
    g = _load(golden_dir, name)
    sd = unet_oracle.seeded_state_dict(seed=0, attention=attention)
    assert len(sd) == int(g["n_keys"]) == (256 if attention else 232)
    n_params = sum(v.numel() for k, v in sd.items()
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_params == int(g["n_params"]) == (3453012 if attention else 2645390)
    with torch.no_grad():
        logits, atts = unet_oracle.unet_forward(sd, torch.from_numpy(g["x"]), attention=attention)
    assert np.abs(logits.numpy() - g["logits"]).max() < 2e-5
    assert len(atts) == (6 if attention else 0)
    for i, a in enumerate(atts):  # coarsest first
        assert a.shape == g[f"att{i}"].shape
        assert np.abs(a.numpy() - g[f"att{i}"]).max() < 2e-6


def test_unet_train_step_matches_reference(golden_dir):
    g = _load(golden_dir, "unet_train_step.npz")
    sd = unet_oracle.seeded_state_dict(seed=1, attention=True)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    logits, atts = unet_oracle.unet_forward(params, x, attention=True, training=True)
    loss = loss_oracle.dice_spvpa_loss(logits, atts, y)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert np.abs(logits.detach().numpy() - g["logits"]).max() < 1e-4
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    for n, ref in zip(names, g["grad_norms"]):
        got = params[n].grad.double().norm().item()
        assert abs(got - ref) <= 2e-3 * max(ref, 1e-6) + 1e-7, n
    for key in g.files:
        if key.startswith("grad::"):
            n = key[len("grad::"):]
            ref = g[key]
            assert np.abs(params[n].grad.numpy() - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-7, n
    # running statistics update (momentum 0.1, unbiased variance)
    assert np.abs(params["model.0.conv.unit0.norm.running_mean"].detach().numpy()
                  - g["rm::model.0.conv.unit0.norm.running_mean"]).max() < 1e-5
    assert np.abs(params["model.0.conv.unit0.norm.running_var"].detach().numpy()
                  - g["rv::model.0.conv.unit0.norm.running_var"]).max() < 1e-5


@pytest.mark.parametrize("case", ["ellipsoid", "empty", "full"])
def test_loss_matches_reference(golden_dir, case):
    g = _load(golden_dir, "dice_spvpa_loss.npz")
    x = torch.from_numpy(g[case + "_x"]).requires_grad_(True)
    y = torch.from_numpy(g[case + "_y"])
    atts = [torch.from_numpy(g[f"{case}_att{i}"]).requires_grad_(True) for i in range(6)]
    for a, h in ((1, 1), (1, 0), (0, 1), (0, 0)):
        loss = loss_oracle.dice_spvpa_loss(x, atts, y, supervised_attention=bool(a), hardness_weighting=bool(h))
        assert abs(loss.item() - float(g[f"{case}_a{a}h{h}_loss"])) < 2e-6, (case, a, h)
    loss = loss_oracle.dice_spvpa_loss(x, atts, y)
    loss.backward()
    assert np.abs(x.grad.numpy() - g[case + "_gx"]).max() <= 1e-4 * np.abs(g[case + "_gx"]).max() + 1e-12
    for i, a in enumerate(atts):
        ref = g[f"{case}_gatt{i}"]
        assert np.abs(a.grad.numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-12

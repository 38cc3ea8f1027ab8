"""Host plumbing on CPU (BASELINE config 0): NIfTI I/O, transforms, VSparams train + inference recipe on
tiny synthetic volumes.  No GPU, no native kernels: the torch containers of the drop-in modules run."""
import argparse
import os
import sys

import numpy as np
import torch

from vs_seg_b200 import dataio


def test_nifti_roundtrip_and_orientation(tmp_path):
    rng = np.random.RandomState(0)
    vol = rng.standard_normal((5, 6, 7)).astype(np.float32)
    aff = np.array([[-0.5, 0, 0, 10], [0, 0.5, 0, -3], [0, 0, 2.0, 4], [0, 0, 0, 1.0]])  # LAS: x flipped
    p = str(tmp_path / "a.nii.gz")
    dataio.write_nifti(p, vol, aff)
    got, got_aff = dataio.read_nifti(p)
    assert got.dtype == np.float32 and np.array_equal(got, vol) and np.allclose(got_aff, aff)
    d = dataio.Compose([dataio.LoadNiftid(keys=["image"]), dataio.AddChanneld(keys=["image"]),
                        dataio.Orientationd(keys=["image"], axcodes="RAS")])({"image": p})
    assert np.array_equal(d["image"][0], vol[::-1])  # flipped to R
    assert d["image_meta_dict"]["affine"][0, 0] > 0
    lab = (vol > 0).astype(np.uint8)
    dataio.write_nifti(str(tmp_path / "l.nii"), lab, aff)
    got_l, _ = dataio.read_nifti(str(tmp_path / "l.nii"))
    assert got_l.dtype == np.uint8 and np.array_equal(got_l, lab)


def test_transforms_pad_crop_flip_cache():
    item = {"image": np.arange(2 * 3 * 4, dtype=np.float32).reshape(1, 2, 3, 4), "label": np.ones((1, 2, 3, 4), np.float32)}
    pad = dataio.SpatialPadd(keys=["image", "label"], spatial_size=[5, 3, 8])(item)
    assert pad["image"].shape == (1, 5, 3, 8) and pad["label"].sum() == 24
    assert pad["image"][0, 1, 0, 2] == 0.0  # symmetric: 1 before / 2 after in x, 2/2 in z
    crop = dataio.RandSpatialCropd(keys=["image", "label"], roi_size=[2, 2, 2]).set_random_state(seed=1)(pad)
    assert crop["image"].shape == (1, 2, 2, 2)
    norm = dataio.NormalizeIntensityd(keys=["image"])(item)["image"]
    assert abs(norm.mean()) < 1e-6 and abs(norm.std() - 1) < 1e-6
    chain = dataio.Compose([dataio.NormalizeIntensityd(keys=["image"]),
                            dataio.RandFlipd(keys=["image", "label"], prob=1.0, spatial_axis=0),
                            dataio.ToTensord(keys=["image", "label"])])
    ds = dataio.CacheDataset([item], chain)
    assert chain.first_random_index() == 1
    out = ds[0]
    assert torch.is_tensor(out["image"]) and torch.equal(out["image"], torch.from_numpy(norm[:, ::-1].copy()))


def test_vsparams_train_and_inference_recipe_on_cpu(tmp_path, monkeypatch):
    """The call order of VS_train.py / VS_inference.py on 6 synthetic 48x56x12 cases padded/cropped to 64x64x16, 2 epochs."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.chdir(root)
    from params.VSparams import VSparams
    data_root = str(tmp_path / "VS_defaced") + "/"
    monkeypatch.setattr(sys, "argv", ["VS_train.py", "--debug", "--dataset", "T1", "--device", "cpu", "--synthetic",
                                      "--num_epochs", "2", "--data_root", data_root])
    p = VSparams(argparse.ArgumentParser())
    p.num_workers = 0
    p.synthetic_shape = (48, 56, 12)   # smaller than the crop: exercises SpatialPadd and the inferer's padding
    p.pad_crop_shape = p.pad_crop_shape_test = p.sliding_window_inferer_roi_size = [64, 64, 16]
    p.create_results_folders()
    p.set_up_logger("training_log.txt")
    p.log_parameters()
    train_files, val_files, test_files = p.load_T1_or_T2_data()
    assert (len(train_files), len(val_files), len(test_files)) == (2, 2, 2)
    train_t, val_t, test_t = p.get_transforms()
    dataio.set_determinism(seed=0)
    p.check_transforms_on_first_validation_image_and_label(val_files, val_t)
    train_loader = p.cache_transformed_train_data(train_files, train_t)
    val_loader = p.cache_transformed_val_data(val_files, val_t)
    model = p.set_and_get_model()
    assert len(model.state_dict()) == 256
    losses, metrics = p.run_training_algorithm(model, p.set_and_get_loss_function(), p.set_and_get_optimizer(model),
                                               train_loader, val_loader)
    assert len(losses) == 2 and len(metrics) == 1 and all(np.isfinite(losses))
    p.plot_loss_curve_and_mean_dice(losses, metrics)
    for f in ("best_metric_model.pth", "last_epoch_model.pth"):
        assert os.path.isfile(os.path.join(p.model_path, f))
    # inference recipe
    test_loader = p.cache_transformed_test_data(test_files, test_t)
    model = p.load_trained_state_of_model(p.set_and_get_model())
    scores = p.run_inference(model, test_loader)
    assert scores.shape == (2,) and np.all((scores >= 0) & (scores <= 1))
    out = os.path.join(p.results_folder_path, "inferred_segmentations_nifti", "vs_gk_202", "vs_gk_seg_refT1",
                       "vs_gk_seg_refT1.nii.gz")
    assert os.path.isfile(out)
    seg, _ = dataio.read_nifti(out)
    assert seg.shape == (48, 56, 12) and set(np.unique(seg)) <= {0, 1}
    # figures of the reference (VSparams.py:530-545, :596-616), drawn with PIL
    for f in ("epoch_average_loss_and_val_mean_dice.png", "best_model_output_dice_score_histogram.png",
              "best_model_output_val0.png", "best_model_output_val1.png", "check_validation_image_and_label.png"):
        assert os.path.getsize(os.path.join(p.figures_path, f)) > 0
    assert os.path.isfile(os.path.join(root, "params", "split_TCIA.csv"))   # the default of --split
    for h in list(p.logger.handlers):
        p.logger.removeHandler(h)


def test_fused_adam_on_cpu_is_torch_adam():
    """CPU parameters (the reference's --debug plumbing run) go through torch.optim.Adam unchanged, including the
    learning-rate halving the training loop does through optimizer.param_groups (reference VSparams.py:517-523)."""
    import torch
    from vs_seg_b200.optim import FusedAdam
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa, ob = FusedAdam(a, lr=1e-2, weight_decay=1e-3), torch.optim.Adam(b, lr=1e-2, weight_decay=1e-3)
    for it in range(5):
        for o, ps in ((oa, a), (ob, b)):
            o.zero_grad()
            sum((p * (i + 1 + it)).sum() for i, p in enumerate(ps)).backward()
            if it == 2:
                for g in o.param_groups:
                    g["lr"] = g["lr"] / 2
            o.step()
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    assert oa.flat_grads() == []


def test_fused_adam_state_dict_round_trip_on_cpu():
    """state_dict / load_state_dict carry the moments and step counters (a resumed run must not restart at zero)."""
    from vs_seg_b200.optim import FusedAdam
    torch.manual_seed(0)
    w1, w2 = torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(5, 3))
    w2.data.copy_(w1.data)
    a, b = FusedAdam([w1], lr=1e-2, weight_decay=1e-3), FusedAdam([w2], lr=1e-2, weight_decay=1e-3)
    g = torch.randn(4, 5, 3)
    for i in range(2):
        w1.grad = g[i].clone()
        a.step()
    import copy
    sd = copy.deepcopy(a.state_dict())   # as torch.save / torch.load would (state_dict() hands out references)
    assert sd["param_groups"][0]["step"] == 2 and "fused" in sd and "cpu" in sd
    w2.data.copy_(w1.data)
    b.load_state_dict(sd)
    for i in range(2, 4):
        w1.grad, w2.grad = g[i].clone(), g[i].clone()
        a.step()
        b.step()
    assert torch.equal(w1.data, w2.data)


def test_nifti_saver_restores_the_original_orientation(tmp_path):
    """NiftiSaver must undo Orientationd(RAS) before writing under the file's original affine (reference
    VSparams.py:582-594 passes affine AND original_affine to MONAI's NiftiSaver): LPS-ordered files (negative
    diagonal, as Slicer exports the VS data) and permuted axes round-trip voxel for voxel."""
    import numpy as np
    from vs_seg_b200 import dataio
    rng = np.random.RandomState(0)
    vol = (rng.rand(8, 6, 4) > 0.5).astype(np.uint8)
    affines = [np.array([[-0.4, 0, 0, 50.], [0, -0.4, 0, 60.], [0, 0, 1.5, -20.], [0, 0, 0, 1]]),
               np.array([[0, -0.5, 0, 10.], [0.7, 0, 0, 5.], [0, 0, -2.0, 3.], [0, 0, 0, 1]]),
               np.diag([1.0, 1.0, 1.0, 1.0])]
    for k, aff in enumerate(affines):
        src = str(tmp_path / f"case{k}.nii.gz")
        dataio.write_nifti(src, vol, aff)
        d = dataio.LoadNiftid(keys=["label"])({"label": src})
        d = dataio.Orientationd(keys=["label"])(dataio.AddChanneld(keys=["label"])(d))
        ras = d["label_meta_dict"]["affine"]
        assert all(ras[i, i] > 0 for i in range(3))                     # the network sees RAS data
        out = dataio.NiftiSaver(output_dir=str(tmp_path / "out"), output_postfix="").save(
            d["label"], meta_data=d["label_meta_dict"])
        back, aff_back = dataio.read_nifti(out)
        assert np.array_equal(back, vol)
        assert np.allclose(aff_back, aff, atol=1e-5)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) prints ONE JSON line with the keys the driver
    reads; it loads none of the repo's native code and needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("VS_inference sliding-window 384x384x160")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0

"""Oracle: MONAI 0.4.0 ``sliding_window_inference`` restated (test infrastructure only).

PARITY UNPINNED against MONAI itself: ``monai==0.4.0`` (requirements.txt:7 of the
reference) is a third-party dependency that is absent from /root/reference and cannot be
installed offline.  This file restates its published algorithm
(monai/inferers/utils.py::sliding_window_inference, monai/data/utils.py::
{dense_patch_slices, compute_importance_map}, monai/networks/layers/convutils.py::gaussian_1d
with approx="erf") and is anchored on the reference's call site
/root/reference/params/VSparams.py:568-574 (overlap 0.25 default, mode="gaussian",
sigma_scale 0.125, constant padding with 0).
"""
from __future__ import annotations

import itertools
import math

import torch
import torch.nn.functional as F


def scan_interval(image_size, roi_size, overlap):
    """MONAI ``_get_scan_interval``."""
    out = []
    for img, roi in zip(image_size, roi_size):
        if roi == img:
            out.append(int(roi))
        else:
            iv = int(roi * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_starts(image_size, roi_size, interval):
    """MONAI ``dense_patch_slices``: start corners, first dim slowest / last dim fastest."""
    per_dim = []
    for img, roi, iv in zip(image_size, roi_size, interval):
        if iv == 0:
            n = 1
        else:
            num = int(math.ceil(float(img) / iv))
            first = next((d for d in range(num) if d * iv + roi >= img), None)
            n = first + 1 if first is not None else 1
        starts = []
        for i in range(n):
            s = i * iv
            s -= max(s + roi - img, 0)
            starts.append(s)
        per_dim.append(starts)
    return list(itertools.product(*per_dim))


def gaussian_1d_erf(sigma, truncated=4.0):
    """MONAI 0.4.0 ``gaussian_1d(sigma, truncated, approx='erf')`` (un-normalised)."""
    tail = int(max(float(sigma) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float)
    t = 0.70710678 / abs(float(sigma))
    out = 0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())
    return out.clamp(min=0)


def importance_map(roi_size, mode="gaussian", sigma_scale=0.125):
    """MONAI ``compute_importance_map``: delta at roi//2 blurred by a separable Gaussian
    (zero padded 'same' convolution), normalised by its max, clamped at its min non-zero."""
    if mode == "constant":
        return torch.ones(tuple(roi_size), dtype=torch.float)
    m = torch.zeros(tuple(roi_size), dtype=torch.float)
    m[tuple(r // 2 for r in roi_size)] = 1
    m = m[None, None]
    nd = len(roi_size)
    for d, r in enumerate(roi_size):
        k = gaussian_1d_erf(r * sigma_scale)
        shape = [1, 1] + [1] * nd
        shape[2 + d] = k.numel()
        pad = [0] * nd
        pad[d] = (k.numel() - 1) // 2
        conv = (F.conv1d, F.conv2d, F.conv3d)[nd - 1]
        m = conv(m, k.reshape(shape), padding=pad)
    m = m[0, 0]
    m = m / m.max()
    nz_min = m[m != 0].min().item()
    return torch.clamp(m, min=nz_min).float()


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25,
                             mode="constant", sigma_scale=0.125, padding_mode="constant", cval=0.0):
    """MONAI 0.4.0 ``sliding_window_inference`` for [B,C,*spatial] inputs."""
    nd = inputs.dim() - 2
    assert 0 <= overlap < 1
    image_size_ = list(inputs.shape[2:])
    batch = inputs.shape[0]
    roi_size = tuple(int(r) for r in roi_size)
    image_size = tuple(max(image_size_[i], roi_size[i]) for i in range(nd))
    pad_size = []
    for k in range(inputs.dim() - 1, 1, -1):
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    inputs = F.pad(inputs, pad=pad_size, mode=padding_mode, value=cval)
    interval = scan_interval(image_size, roi_size, overlap)
    starts = window_starts(image_size, roi_size, interval)
    num_win = len(starts)
    total = num_win * batch
    imap = importance_map(roi_size, mode=mode, sigma_scale=sigma_scale).to(inputs.device)
    out = cnt = None
    for g0 in range(0, total, sw_batch_size):
        idxs = range(g0, min(g0 + sw_batch_size, total))
        sl = []
        for idx in idxs:
            b, w = idx // num_win, idx % num_win
            sl.append((slice(b, b + 1), slice(None)) +
                      tuple(slice(s, s + r) for s, r in zip(starts[w], roi_size)))
        win = torch.cat([inputs[s] for s in sl])
        prob = predictor(win)
        if out is None:
            shape = [batch, prob.shape[1]] + list(image_size)
            out = torch.zeros(shape, dtype=torch.float32, device=inputs.device)
            cnt = torch.zeros(shape, dtype=torch.float32, device=inputs.device)
        for j, s in enumerate(sl):
            out[s] += imap * prob[j]
            cnt[s] += imap
    out = out / cnt
    final = [slice(None), slice(None)]
    for sp in range(nd):
        lo = pad_size[(nd - 1 - sp) * 2]
        final.append(slice(lo, lo + image_size_[sp]))
    return out[tuple(final)]

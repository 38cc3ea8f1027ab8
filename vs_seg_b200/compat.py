"""Names from MONAI 0.4.0 that appear in the reference's public call sites
(/root/reference/params/VSparams.py:343-374 passes ``norm=Norm.BATCH``; the network files use
Act/Norm/Dropout factories, ``same_padding`` and ``SkipConnection``).  MONAI is not a dependency
of this package; these are minimal equivalents with the same spelling and behaviour."""
from __future__ import annotations

from enum import Enum

import numpy as np
import torch
import torch.nn as nn


class _Names:
    def __init__(self, table):
        self._table = {k.upper(): v for k, v in table.items()}
        for k in self._table:
            setattr(self, k, k)

    def __getitem__(self, key):
        if isinstance(key, tuple):
            name, dim = key
            return self._table[str(name).upper()][int(dim)]
        return self._table[str(key).upper()]


Conv = _Names({"CONV": {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d},
               "CONVTRANS": {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}})
Norm = _Names({"BATCH": {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d},
               "INSTANCE": {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}})
Dropout = _Names({"DROPOUT": {1: nn.Dropout, 2: nn.Dropout2d, 3: nn.Dropout3d}})
Act = _Names({"PRELU": nn.PReLU, "RELU": nn.ReLU, "SIGMOID": nn.Sigmoid})


def split_args(args):
    """'name' -> ('name', {}); ('name', {...}) -> as is."""
    if isinstance(args, str):
        return args, {}
    name, kw = args
    return name, dict(kw)


def same_padding(kernel_size, dilation=1):
    k = np.atleast_1d(kernel_size)
    d = np.atleast_1d(dilation)
    if np.any((k - 1) * d % 2 == 1):
        raise NotImplementedError(f"Same padding not available for kernel_size={k}, dilation={d}.")
    p = tuple(int(v) for v in (k - 1) / 2 * d)
    return p if len(p) > 1 else p[0]


class SkipConnection(nn.Module):
    """cat([x, submodule(x)], dim) — the child is registered as ``submodule`` (state_dict keys)."""

    def __init__(self, submodule, cat_dim=1):
        super().__init__()
        self.submodule = submodule
        self.cat_dim = cat_dim

    def forward(self, x):
        return torch.cat([x, self.submodule(x)], self.cat_dim)


class LossReduction(Enum):
    NONE = "none"
    MEAN = "mean"
    SUM = "sum"


def one_hot(labels, num_classes, dtype=torch.float, dim=1):
    if labels.dim() < dim + 1:
        raise AssertionError("labels should have a channel dimension")
    shape = list(labels.shape)
    if shape[dim] != 1:
        raise AssertionError("labels should have a channel with length equals to one.")
    shape[dim] = num_classes
    out = torch.zeros(size=shape, dtype=dtype, device=labels.device)
    return out.scatter_(dim=dim, index=labels.long(), value=1)

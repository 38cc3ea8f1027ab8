import torch.nn as nn


class _Factory:
    def __init__(self, table):
        self._t = table
        for k in table:
            setattr(self, k.upper(), k.upper())

    def __getitem__(self, key):
        if isinstance(key, tuple):
            name, dim = key
            return self._t[name.lower()][dim]
        return self._t[key.lower()]


Conv = _Factory({"conv": {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d},
                 "convtrans": {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}})
Norm = _Factory({"batch": {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d},
                 "instance": {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}})
Dropout = _Factory({"dropout": {1: nn.Dropout, 2: nn.Dropout2d, 3: nn.Dropout3d}})
Act = _Factory({"prelu": nn.PReLU, "relu": nn.ReLU, "sigmoid": nn.Sigmoid})


def split_args(args):
    if isinstance(args, str):
        return args, {}
    name, a = args
    return name, a

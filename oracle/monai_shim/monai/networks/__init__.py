import torch


def one_hot(labels, num_classes, dtype=torch.float, dim=1):
    shape = list(labels.shape)
    assert shape[dim] == 1
    shape[dim] = num_classes
    out = torch.zeros(size=shape, dtype=dtype, device=labels.device)
    return out.scatter_(dim=dim, index=labels.long(), value=1)

import numpy as np


def same_padding(kernel_size, dilation=1):
    k = np.atleast_1d(kernel_size)
    d = np.atleast_1d(dilation)
    p = (k - 1) / 2 * d
    p = tuple(int(v) for v in p)
    return p if len(p) > 1 else p[0]

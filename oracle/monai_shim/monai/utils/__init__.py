from enum import Enum

from .aliases import alias


def export(modname):
    def deco(obj):
        return obj
    return deco


class LossReduction(Enum):
    NONE = "none"
    MEAN = "mean"
    SUM = "sum"


class Weight(Enum):
    SQUARE = "square"
    SIMPLE = "simple"
    UNIFORM = "uniform"

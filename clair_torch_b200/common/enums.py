"""Boundary constants of the hot path (values match clair_torch/common/enums.py:10-17)."""
from enum import Enum, auto


class InterpMode(Enum):
    """How ICRFModelBase.forward evaluates the (C, L) table."""
    LOOKUP = auto()   # nearest sample, round-half-even, no gradient
    LINEAR = auto()   # two-tap interpolation (default everywhere in the reference)
    CATMULL = auto()  # four-tap Catmull-Rom


class VarianceMode(Enum):
    """Normalisation of the running second moment (clair_torch/common/enums.py, common/statistics.py:153-173)."""
    POPULATION = auto()            # M2 / W
    SAMPLE_FREQUENCY = auto()      # M2 / (W - 1)
    RELIABILITY_WEIGHTS = auto()   # M2 / (W - W2 / W)

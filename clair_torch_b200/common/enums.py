"""Boundary constants of the hot path (values match clair_torch/common/enums.py:10-17)."""
from enum import Enum, auto


class InterpMode(Enum):
    """How ICRFModelBase.forward evaluates the (C, L) table."""
    LOOKUP = auto()   # nearest sample, round-half-even, no gradient
    LINEAR = auto()   # two-tap interpolation (default everywhere in the reference)
    CATMULL = auto()  # four-tap Catmull-Rom: SURVEY.md §8(f) rank 4, not built yet

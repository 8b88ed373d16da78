"""Concrete ICRF models (clair_torch/models/icrf_model.py)."""
from typing import Optional

import torch
from torch import nn

from ..common.enums import InterpMode
from .base import ICRFModelBase


class ICRFModelDirect(ICRFModelBase):
    """Every table sample is a parameter: C parameter vectors of length n_points
    (clair_torch/models/icrf_model.py:89-127).

    As in the reference the parameters are ALWAYS initialised to linspace(0,1)**initial_power, even when an
    `icrf` table is passed, and `_icrf` only becomes a function of them at the first update_icrf() — so the
    first optimiser step of train_icrf moves nothing.
    """

    def __init__(self, n_points: Optional[int] = 256, channels: Optional[int] = 3,
                 interpolation_mode: InterpMode = InterpMode.LINEAR, initial_power: float = 2.5,
                 icrf: Optional[torch.Tensor] = None):
        super().__init__(n_points, channels, interpolation_mode, initial_power, icrf)
        self.direct_params = nn.ParameterList([
            nn.Parameter(torch.linspace(0, 1, n_points) ** initial_power) for _ in range(channels)
        ])

    def channel_params(self, c: int):
        return [self.direct_params[c]]

    def update_icrf(self):
        self._icrf = torch.stack([p for p in self.direct_params], dim=0)

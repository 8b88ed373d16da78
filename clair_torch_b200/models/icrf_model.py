"""Concrete ICRF models (clair_torch/models/icrf_model.py)."""
from typing import Optional

import torch
from torch import nn

from ..common.enums import InterpMode
from .base import ICRFModelBase
from ..common.errors import ArgumentTypeError


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


class ICRFModelPCA(ICRFModelBase):
    """Power-law base curve plus a principal-component expansion, one exponent and `num_components` coefficients per
    channel (clair_torch/models/icrf_model.py:14-86).  `pca_basis` has shape (n_points, num_components, channels).

    Deviation from the reference, on purpose: its update_icrf stores the table as (L, C) while forward expects (C, L)
    (SURVEY.md Q2), so after the first update its forward reads the wrong axis; here the table stays (C, L).
    """

    def __init__(self, pca_basis: torch.Tensor, interpolation_mode: InterpMode = InterpMode.LINEAR,
                 initial_power: float = 2.5, icrf: Optional[torch.Tensor] = None) -> None:
        if not isinstance(pca_basis, torch.Tensor) or pca_basis.dim() != 3:
            raise ArgumentTypeError("pca_basis must be a (n_points, num_components, channels) tensor")
        n_points, num_components, channels = pca_basis.shape
        super().__init__(n_points, channels, interpolation_mode, initial_power, icrf)
        self.p = nn.ParameterList([nn.Parameter(torch.tensor(2.0)) for _ in range(channels)])
        self.coefficients = nn.ParameterList([nn.Parameter(torch.zeros(num_components)) for _ in range(channels)])
        self.register_buffer("pca_basis", pca_basis)
        self.register_buffer("x_values", torch.linspace(0, 1, n_points))

    def channel_params(self, c: int):
        return [self.p[c], self.coefficients[c]]

    def update_icrf(self):
        p = torch.stack(list(self.p))                                         # (C,)
        coeff = torch.stack(list(self.coefficients))                           # (C, K)
        base = self.x_values.clamp(min=1e-6).unsqueeze(1).pow(p.unsqueeze(0))  # (L, C), icrf_model.py:77-80
        pca = (self.pca_basis * coeff.T.unsqueeze(0)).sum(dim=1)               # (L, C), :83
        self._icrf = (base + pca).transpose(0, 1).contiguous()                 # kept (C, L)

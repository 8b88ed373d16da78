"""ICRF model base class: same constructor, properties and forward contract as
clair_torch/models/base.py:17-136, with the table evaluation done by the sm_100a kernels."""
from abc import ABC, abstractmethod
from typing import Optional

import torch
from torch import nn

from .. import _native, kernels
from ..common.enums import InterpMode
from ..common.errors import ArgumentTypeError


class _TableFn(torch.autograd.Function):
    """f = table interpolation (LINEAR models/base.py:160-182, CATMULL :184-226) with both of its autograd edges:
    d/d image = f'(x) (elementwise) and d/d table = tap scatter (the index_put of :176 / :219).  LOOKUP (:138-158) is a
    gather: it has the table edge only (callers hand it a detached image)."""

    @staticmethod
    def forward(ctx, image, table, interp_mode):
        if interp_mode == _native.INTERP_LOOKUP:
            y, dydx = kernels.icrf_forward(image, table, interp_mode), None
        else:
            y, dydx = kernels.icrf_forward(image, table, interp_mode, want_derivative=True)
        ctx.save_for_backward(image.detach(), dydx)
        ctx.table_shape = tuple(table.shape)
        ctx.table_dtype = table.dtype
        ctx.interp_mode = interp_mode
        return y

    @staticmethod
    def backward(ctx, grad_out):
        image, dydx = ctx.saved_tensors
        g_image = g_table = None
        if ctx.needs_input_grad[0] and dydx is not None:
            g_image = grad_out * dydx
        if ctx.needs_input_grad[1]:
            c, lut = ctx.table_shape
            g_table = kernels.icrf_backward_theta(image, grad_out.to(torch.float32), c, lut,
                                                  interp_mode=ctx.interp_mode).to(ctx.table_dtype)
        return g_image, g_table, None


class ICRFModelBase(nn.Module, ABC):
    """Inverse camera response function as a (C, L) table over [0, 1]."""

    def __init__(self, n_points: Optional[int] = 256, channels: Optional[int] = 3,
                 interpolation_mode: InterpMode = InterpMode.LINEAR, initial_power: float = 2.5,
                 icrf: Optional[torch.Tensor] = None):
        super().__init__()
        if not isinstance(interpolation_mode, InterpMode):
            raise ArgumentTypeError(f"interpolation_mode must be an InterpMode, got {type(interpolation_mode)}")
        if icrf is not None:
            if not isinstance(icrf, torch.Tensor):
                raise ArgumentTypeError("icrf must be a torch.Tensor")
            channels, n_points = icrf.shape          # overrides n_points / channels, base.py:59-60
        self._channels = channels
        self._initial_power = initial_power
        self._n_points = n_points
        self.register_buffer("_x_axis_datapoints", torch.linspace(0, 1, n_points))
        if icrf is None:
            icrf = self._initialize_default_icrf()
        self.register_buffer("_icrf", icrf)
        self.interpolation_mode = interpolation_mode
        if interpolation_mode not in (InterpMode.LOOKUP, InterpMode.LINEAR, InterpMode.CATMULL):
            raise ValueError(f"Unknown interpolation mode {interpolation_mode}")

    icrf = property(lambda self: self._icrf)
    channels = property(lambda self: self._channels)
    n_points = property(lambda self: self._n_points)
    initial_power = property(lambda self: self._initial_power)
    x_axis_datapoints = property(lambda self: self._x_axis_datapoints)

    @abstractmethod
    def channel_params(self, c: int) -> list[nn.Parameter]:
        """Optimisation parameters of channel c (fed to one torch optimiser per channel)."""

    @abstractmethod
    def update_icrf(self) -> None:
        """Rebuild self._icrf from the parameters."""

    def _initialize_default_icrf(self) -> torch.Tensor:
        x = torch.linspace(0, 1, self.n_points).unsqueeze(1).repeat(1, self.channels)
        return torch.transpose(x ** self.initial_power, 0, 1)

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        """(N, C, H, W) fp32 image stack on a CUDA device -> linearised stack of the same shape."""
        if self.interpolation_mode is InterpMode.LINEAR:
            return _TableFn.apply(image, self._icrf, _native.INTERP_LINEAR)
        if self.interpolation_mode is InterpMode.LOOKUP:
            return _TableFn.apply(image.detach(), self._icrf, _native.INTERP_LOOKUP)
        # the reference's own CATMULL forward only runs on the CPU (models/base.py:218 builds arange without a device)
        return _TableFn.apply(image, self._icrf, _native.INTERP_CATMULL)

    def plot_icrf(self) -> None:
        """Live plotting lives in the reference's visualization package and is out of scope here (no-op)."""

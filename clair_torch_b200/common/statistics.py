"""Weighted batched online mean / variance over the frame dimension, kernel-backed
(clair_torch/common/statistics.py:13-259).

Only what the reference's own drivers use is supported: `dim=0` over `(N, C, H, W)` fp32 CUDA batches.  The running
state lives in four `(C, H, W)` fp32 device buffers and every batch is one pass of `clair_frame_stats_update`.
`WBOMean` as used inside compute_hdr_image is NOT this class — there the merge is fused into the HDR kernel.
"""
import ctypes
from typing import Optional

import torch

from .. import _native
from .enums import VarianceMode
from .errors import ArgumentTypeError


class WBOMean:
    def __init__(self, dim: int | tuple[int, ...] = 0):
        if isinstance(dim, int):
            dim = (dim,)
        else:
            raise ArgumentTypeError(f"Expected dim as int or tuple of int, got {type(dim)}")     # as statistics.py:24-27
        if dim != (0,):
            raise NotImplementedError("the kernel-backed running statistics reduce over the frame dimension (dim=0)")
        self._dim = dim
        self._mean = self._m2 = self._sum_of_weights = self._sum_of_squared_weights = None
        self._batches = 0

    mean = property(lambda self: 0.0 if self._mean is None else self._mean.unsqueeze(0))
    sum_of_weights = property(lambda self: 0.0 if self._sum_of_weights is None else self._sum_of_weights.unsqueeze(0))
    dim = property(lambda self: self._dim)

    def internal_detach(self, *, in_place: bool = True):
        """The state never carries an autograd graph here; kept for call compatibility."""

    def _update(self, batch_values: torch.Tensor, batch_weights: Optional[torch.Tensor], table=None,
                interp_mode: int = _native.INTERP_LINEAR):
        from ..kernels import _ptr, _stack, _stream, _table
        lib = _native.load()
        val = _stack(batch_values, "batch_values")
        wts = None if batch_weights is None else _stack(batch_weights, "batch_weights")
        if wts is not None and wts.shape != val.shape:
            raise ValueError("batch_weights must have the shape of batch_values")
        n, c, h, w = val.shape
        th = _table(table, val.device, c)
        if self._mean is None:
            make = lambda: torch.empty((c, h, w), dtype=torch.float32, device=val.device)
            self._mean, self._m2, self._sum_of_weights, self._sum_of_squared_weights = make(), make(), make(), make()
        with torch.cuda.device(val.device):
            rc = lib.clair_frame_stats_update(_ptr(val), _ptr(wts), _ptr(th), n, c, h * w, 0 if th is None else th.shape[1],
                                              int(interp_mode), None, _ptr(self._mean), _ptr(self._m2), _ptr(self._sum_of_weights),
                                              _ptr(self._sum_of_squared_weights), int(self._batches == 0), _stream(val.device))
        _native.check(rc, "clair_frame_stats_update")
        self._batches += 1

    def update_values(self, batch_values: torch.Tensor, batch_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._update(batch_values, batch_weights)
        return self.mean


class WBOMeanVar(WBOMean):
    def __init__(self, dim: int | tuple[int, ...] = 0, variance_mode: VarianceMode = VarianceMode.RELIABILITY_WEIGHTS):
        super().__init__(dim=dim)
        if not isinstance(variance_mode, VarianceMode):
            raise ValueError(f"Unknown variance mode {variance_mode}")
        self._variance_mode = variance_mode

    sum_of_squared_weights = property(
        lambda self: 0.0 if self._sum_of_squared_weights is None else self._sum_of_squared_weights.unsqueeze(0))
    m2 = property(lambda self: 0.0 if self._m2 is None else self._m2.unsqueeze(0))

    def variance(self):
        w, w2 = self.sum_of_weights, self.sum_of_squared_weights
        if self._variance_mode is VarianceMode.SAMPLE_FREQUENCY:
            return self.m2 * (1 / (w - 1))
        if self._variance_mode is VarianceMode.RELIABILITY_WEIGHTS:
            return self.m2 * (1 / (w - w2 / w))
        return self.m2 * (1 / w)

    def update_values(self, batch_values: torch.Tensor, batch_weights: Optional[torch.Tensor] = None, *, table=None,
                      interp_mode: int = _native.INTERP_LINEAR):
        """Returns (mean, m2) like the reference.  `table` (C, L) linearises the frames inside the same pass
        (`interp_mode`: the InterpMode of the model the table belongs to)."""
        self._update(batch_values, batch_weights, table, interp_mode)
        return self.mean, self.m2

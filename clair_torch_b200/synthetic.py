"""Synthetic exposure stacks of the shapes BASELINE.json names.

Scene radiance E = exp(U(-7, 0)) per element, exposures t_k = t0 * 2^k, camera response v = clip(4 E t_k / t_max, 0, 1)^(1/2.2),
quantised to `bits` and normalised the way the reference's CastTo + Normalize transforms do it
(fl32(code) / fl32(maxval), clair_torch/common/transforms.py:107-190), std = 0.05 * value
(MissingStdMode.MULTIPLIER, clair_torch/datasets/base.py:132-133).
"""
from typing import Optional

import numpy as np
import torch


def exposure_times(n_frames: int, first: float = 1e-3) -> np.ndarray:
    return (first * 2.0 ** np.arange(n_frames)).astype(np.float64)


def reference_curve(channels: int = 3, lut: int = 256, powers=(2.2, 2.0, 2.4, 1.8, 2.6, 2.1, 2.3, 1.9)) -> torch.Tensor:
    """Fixed (C, L) table with DISTINCT rows so the k-mod-C row striping is exercised."""
    x = torch.linspace(0, 1, lut)
    return torch.stack([x ** powers[c % len(powers)] for c in range(channels)]).to(torch.float32)


def make_stack(n_frames: int, channels: int, height: int, width: int, bits: int = 8, seed: int = 1234,
               device="cpu", std_multiplier: Optional[float] = 0.05, first_exposure: float = 1e-3):
    """Returns (val (N,C,H,W) fp32, std (N,C,H,W) fp32 | None, exposure (N,) float64 numpy)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    scene = torch.exp(torch.rand((channels, height, width), generator=gen, device=dev, dtype=torch.float32) * 7.0 - 7.0)
    t = exposure_times(n_frames, first_exposure)
    maxval = float(2 ** bits - 1)
    val = torch.empty((n_frames, channels, height, width), dtype=torch.float32, device=dev)
    # a tensor divisor: torch's CUDA division by a Python scalar multiplies by its reciprocal, which is not fl32(code / max)
    # for ~3 % of the 16-bit codes; the CPU-made and device-made stacks of one seed must hold the same IEEE quotients
    divisor = torch.tensor(maxval, dtype=torch.float32, device=dev)
    for k in range(n_frames):
        v = torch.clamp(scene * float(4.0 * t[k] / t[-1]), 0.0, 1.0) ** (1.0 / 2.2)
        val[k] = torch.round(v * maxval) / divisor      # fp32 true division, as Normalize does
    std = None if std_multiplier is None else val * std_multiplier
    return val, std, t

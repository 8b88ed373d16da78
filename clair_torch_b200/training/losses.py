"""The eight loss functions re-exported by clair_torch/training/__init__.py:6-14.

All eight are kept as torch functions for call compatibility only.  train_icrf evaluates the four curve penalties and
their gradients on the device (clair_curve_penalties, one kernel), and train_icrf / measure_linearity never materialise
the (P, C, H, W) results of the four per-pixel functions — the fused kernels (clair_pair_stats / clair_pair_grad /
clair_pair_fused) compute the same quantities per pixel in registers.
"""
from typing import Optional

import torch

from ..common.general_functions import weighted_mean_and_std


# ---- curve penalties (training/losses.py:111-190), per-channel or summed ----
def compute_monotonicity_penalty(curve: torch.Tensor, squared=True, per_channel: bool = False) -> torch.Tensor:
    df = curve[:, 1:] - curve[:, :-1]
    falling = (df <= 0).float()
    penalty = (falling * df.pow(2) if squared else falling * (-df)).sum(dim=1)
    return penalty if per_channel else torch.sum(penalty)


def compute_smoothness_penalty(curve: torch.Tensor, per_channel: bool = False) -> torch.Tensor:
    second = curve[:, :-2] - 2 * curve[:, 1:-1] + curve[:, 2:]
    penalty = second.pow(2).sum(dim=1)
    return penalty if per_channel else torch.sum(penalty)


def compute_range_penalty(curve: torch.Tensor, epsilon: float = 1e-6, per_channel: bool = False) -> torch.Tensor:
    penalty = (torch.relu(-curve) + torch.relu(curve - 1)).sum(dim=1)
    return penalty if per_channel else torch.sum(penalty)


def compute_endpoint_penalty(curve: torch.Tensor, per_channel: Optional[bool] = False) -> torch.Tensor:
    if curve.ndim == 1:
        curve = curve.unsqueeze(1)
    if curve.ndim not in (1, 2):
        raise ValueError(f"curve must have 1 or 2 dimensions, got {curve.ndim}")
    penalty = (curve[:, 0] - 0) ** 2 + (curve[:, -1] - 1) ** 2
    return penalty if per_channel else torch.sum(penalty)


# ---- per-pixel functions, compatibility forms ----
def gaussian_value_weights(image: torch.Tensor, scale: Optional[float] = 30.0) -> torch.Tensor:
    """exp(-scale (x - 0.5)^2), training/losses.py:193-205."""
    return torch.exp(-scale * (image - 0.5) ** 2)


def combined_gaussian_pair_weights(image_stack: torch.Tensor, i_idx: torch.Tensor, j_idx: torch.Tensor,
                                   scale: Optional[float] = 10.0) -> torch.Tensor:
    """training/losses.py:208-235."""
    if i_idx.ndim != 1 or j_idx.ndim != 1:
        raise ValueError("i_idx and j_idx must be 1-dimensional")
    g = gaussian_value_weights(image_stack, scale)
    return g[i_idx] + g[j_idx]


def pixelwise_linearity_loss(image_value_stack: torch.Tensor, i_idx: torch.Tensor, j_idx: torch.Tensor,
                             ratio_pairs: torch.Tensor, image_std_stack: Optional[torch.Tensor] = None,
                             use_relative: bool = True):
    """training/losses.py:13-67 (materialises (P, C, H, W))."""
    a, b = image_value_stack[i_idx], image_value_stack[j_idx]
    r = ratio_pairs.view(-1, 1, 1, 1)
    expected = b * r
    diff = a - expected
    safe = expected + 1e-6
    if use_relative:
        diff = diff / safe
    err = None
    if image_std_stack is not None:
        sa, sb = image_std_stack[i_idx], image_std_stack[j_idx]
        if use_relative:
            err = torch.sqrt((sa / safe) ** 2 + ((a * sb) / (safe * b.clamp(min=1e-6))) ** 2 + 1e-6)
        else:
            err = torch.sqrt(sa ** 2 + (r * sb) ** 2)
    return diff.abs(), err


def compute_spatial_linearity_loss(pixelwise_losses: torch.Tensor, pixelwise_errors: Optional[torch.Tensor] = None,
                                   external_weights: Optional[torch.Tensor] = None,
                                   valid_mask: Optional[torch.Tensor] = None, use_uncertainty_weighting: bool = True):
    """training/losses.py:70-108."""
    weights = None
    if pixelwise_errors is not None or external_weights is not None:
        weights = torch.zeros_like(pixelwise_losses)
        if pixelwise_errors is not None and use_uncertainty_weighting:
            weights = weights + 1 / (pixelwise_errors + 1e-6)
        if external_weights is not None:
            weights = weights + external_weights
    mean, std = weighted_mean_and_std(pixelwise_losses, weights=weights, mask=valid_mask, dim=(2, 3))
    err = None
    if pixelwise_errors is not None:
        err, _ = weighted_mean_and_std(pixelwise_errors, mask=valid_mask, dim=(2, 3))
    return mean, std, err

"""Host-side helpers of the pair path.

`get_valid_exposure_pairs` is on the product path (it builds the P-entry pair table the kernels take as an
argument; N <= 64 numbers, host arithmetic).  The other two functions exist for call compatibility with
clair_torch.common: the fused kernels never materialise a (P, C, H, W) mask or weight tensor, so these
tensor-returning forms are NOT used by compute_hdr_image / measure_linearity / train_icrf here.
"""
from typing import Optional

import torch
from .errors import ArgumentTypeError


def get_valid_exposure_pairs(increasing_exposure_values: torch.Tensor,
                             exposure_ratio_threshold: Optional[float] = None):
    """Pairs (i < j) in row-major upper-triangle order with ratio t_i / t_j >= threshold.

    Same outputs as clair_torch/common/general_functions.py:242-272: (i_idx int64, j_idx int64, ratio_pairs in the
    dtype of the exposures), on the device of the exposures.
    """
    if not isinstance(increasing_exposure_values, torch.Tensor):
        raise ArgumentTypeError("increasing_exposure_values must be a torch.Tensor")
    if exposure_ratio_threshold is not None and not isinstance(exposure_ratio_threshold, (int, float)):
        raise ArgumentTypeError("exposure_ratio_threshold must be a float or None")
    t = increasing_exposure_values
    n = t.shape[0]
    i_idx, j_idx = torch.triu_indices(n, n, offset=1)
    i_idx, j_idx = i_idx.to(t.device), j_idx.to(t.device)
    ratio_pairs = t[i_idx] / t[j_idx]
    if exposure_ratio_threshold is not None:
        keep = ratio_pairs >= exposure_ratio_threshold
        i_idx, j_idx, ratio_pairs = i_idx[keep], j_idx[keep], ratio_pairs[keep]
    return i_idx, j_idx, ratio_pairs


def get_pairwise_valid_pixel_mask(image_value_stack: torch.Tensor, i_idx: torch.Tensor, j_idx: torch.Tensor,
                                  image_std_stack: Optional[torch.Tensor] = None, val_lower: float = 0.0,
                                  val_upper: float = 1.0, std_lower: Optional[float] = None,
                                  std_upper: Optional[float] = None) -> torch.Tensor:
    """Compatibility form of clair_torch/common/general_functions.py:276-312 (materialises (P, C, H, W))."""
    if val_lower > val_upper:
        raise ValueError("Lower threshold cannot be a larger value than upper threshold.")
    if std_lower is not None and std_upper is not None and std_lower > std_upper:
        raise ValueError("Lower threshold cannot be a larger value than upper threshold.")
    ok = (image_value_stack >= val_lower) & (image_value_stack <= val_upper)
    if image_std_stack is not None and (std_lower is not None or std_upper is not None):
        ok = ok & (image_std_stack >= std_lower) & (image_std_stack <= std_upper)
    return ok[i_idx] & ok[j_idx]


def weighted_mean_and_std(values: torch.Tensor, weights: Optional[torch.Tensor] = None,
                          mask: Optional[torch.Tensor] = None, dim=None, keepdim: bool = False, eps: float = 1e-8,
                          compute_std: bool = True):
    """Compatibility form of clair_torch/common/general_functions.py:118-178."""
    std = None
    if mask is not None:
        mask = mask.to(dtype=values.dtype)
        values = values * mask
        weights = weights * mask if weights is not None else mask
    if weights is None:
        mean = values.mean(dim=dim, keepdim=True)
        if compute_std:
            std = torch.sqrt(((values - mean) ** 2).mean(dim=dim, keepdim=True))
    else:
        total = weights.sum(dim=dim, keepdim=True).clamp(min=eps)
        mean = (values * weights).sum(dim=dim, keepdim=True) / total
        if compute_std:
            std = torch.sqrt((((values - mean) ** 2) * weights).sum(dim=dim, keepdim=True) / total)
    if not keepdim:
        mean = mean.squeeze(dim) if dim is not None else mean.squeeze()
        if compute_std:
            std = std.squeeze(dim) if dim is not None else std.squeeze()
    return mean, std

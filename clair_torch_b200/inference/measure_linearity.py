"""Linearity measurement driver — call-compatible with clair_torch/inference/measure_linearity.py:17-74."""
from typing import Optional

import torch
from torch.utils.data import DataLoader

from .. import kernels
from ..common.enums import InterpMode
from ..common.general_functions import get_pairwise_valid_pixel_mask, get_valid_exposure_pairs
from ..models.base import ICRFModelBase
from ._common import as_device, linear_table, stage_batch
from ..common.errors import ArgumentTypeError

RATIO_THRESHOLD = 0.2                     # measure_linearity.py:45
VALID_LO, VALID_HI = 1 / 255, 254 / 255   # measure_linearity.py:46


def spatial_statistics(sums: torch.Tensor, with_errors: bool):
    """(P,C,5) float64 sums -> (mean, std, errmean) exactly as weighted_mean_and_std forms them
    (common/general_functions.py:153-171): clamp(min=1e-8) on both denominators."""
    s0, s1, s2, s3, s4 = sums.unbind(dim=-1)
    denom = s0.clamp(min=1e-8)
    mean = s1 / denom
    m2 = (s2 - 2.0 * mean * s1 + mean * mean * s0).clamp(min=0.0)
    std = torch.sqrt(m2 / denom)
    errmean = s3 / s4.clamp(min=1e-8) if with_errors else None
    return mean, std, errmean


def composed_linearity_terms(icrf_model, images, stds, i_idx, j_idx, ratio_pairs, lo, hi, relative: bool, unc_weighting: bool,
                             create_graph: bool = False):
    """LOOKUP / CATMULL models (not the reference default, and CATMULL is CPU-only in the reference: models/base.py:218
    builds its channel index without a device): the reference's own sequence measure_linearity.py:45-72 on the device —
    the model's forward / derivative kernels plus the pairwise algebra as torch ops on materialised (P, C, H, W) tensors.
    Returns (spatial mean, spatial std, mean uncertainty | None) as compute_spatial_linearity_loss does."""
    from ..training.losses import combined_gaussian_pair_weights, compute_spatial_linearity_loss, pixelwise_linearity_loss
    dev = images.device
    i_idx, j_idx = i_idx.to(dev), j_idx.to(dev)
    ratio_pairs = ratio_pairs.to(dev)
    valid = get_pairwise_valid_pixel_mask(images, i_idx, j_idx, stds, val_lower=lo, val_upper=hi)
    gauss = combined_gaussian_pair_weights(images, i_idx, j_idx)
    x = images.detach().requires_grad_(stds is not None)
    with torch.enable_grad():
        linearized = icrf_model(x)
        lin_std = None
        if stds is not None:
            # measure_linearity.py:57-63; a LOOKUP model has no image edge, so this raises as in the reference
            (grads,) = torch.autograd.grad(linearized, x, torch.ones_like(linearized), retain_graph=True)
            lin_std = (grads * stds).abs()
        if not create_graph:
            linearized = linearized.detach()
        loss, err = pixelwise_linearity_loss(linearized, i_idx, j_idx, ratio_pairs, lin_std, relative)
        return compute_spatial_linearity_loss(loss, err, gauss, valid, unc_weighting)


def measure_linearity(dataloader: DataLoader, device, use_uncertainty_weighting: bool = True,
                      use_relative_linearity_loss: bool = True, icrf_model: Optional[ICRFModelBase] = None):
    """(exposure ratios (P,), spatial loss mean (P,C), its std (P,C), its mean uncertainty (P,C) | None), float64.

    Like the reference, only the FIRST batch of the dataloader is measured (the `return` at :74 sits inside the
    loop), so pass the whole stack as one batch.
    """
    if not isinstance(dataloader, DataLoader):
        raise ArgumentTypeError(f"dataloader must be a torch DataLoader, got {type(dataloader)}")
    for flag in (use_uncertainty_weighting, use_relative_linearity_loss):
        if not isinstance(flag, bool):
            raise ArgumentTypeError("use_uncertainty_weighting / use_relative_linearity_loss must be bool")
    dev = as_device(device)
    fused = icrf_model is None or not isinstance(icrf_model, ICRFModelBase) or icrf_model.interpolation_mode is InterpMode.LINEAR
    table = linear_table(icrf_model, dev) if fused else None
    for _, val_batch, std_batch, meta_batch in dataloader:
        images, stds = stage_batch(val_batch, std_batch, dev)
        exposures = meta_batch["exposure_time"]
        i_idx, j_idx, ratio_pairs = get_valid_exposure_pairs(exposures, RATIO_THRESHOLD)
        if not fused:
            mean, std, errmean = composed_linearity_terms(icrf_model, images, stds, i_idx, j_idx, ratio_pairs, VALID_LO, VALID_HI,
                                                          use_relative_linearity_loss, use_uncertainty_weighting)
            return ratio_pairs.to(dev), mean, std, errmean
        sums = kernels.pair_stats(images, stds, i_idx, j_idx, ratio_pairs, table, VALID_LO, VALID_HI,
                                  use_relative_linearity_loss, use_uncertainty_weighting)
        mean, std, errmean = spatial_statistics(sums, stds is not None)
        return ratio_pairs.to(dev), mean, std, errmean
    raise ValueError("the dataloader yielded no batches")

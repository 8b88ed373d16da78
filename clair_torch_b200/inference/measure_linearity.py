"""Linearity measurement driver — call-compatible with clair_torch/inference/measure_linearity.py:17-74."""
from typing import Optional

import torch
from torch.utils.data import DataLoader

from .. import kernels
from ..common.general_functions import get_valid_exposure_pairs
from ..models.base import ICRFModelBase
from ._common import as_device, model_table, stage_batch
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


def measure_linearity(dataloader: DataLoader, device, use_uncertainty_weighting: bool = True,
                      use_relative_linearity_loss: bool = True, icrf_model: Optional[ICRFModelBase] = None, *,
                      code_max: Optional[float] = None):
    """(exposure ratios (P,), spatial loss mean (P,C), its std (P,C), its mean uncertainty (P,C) | None), float64.

    One fused kernel pass over the batch for a model in any InterpMode (LINEAR, LOOKUP, CATMULL) or none; as in the
    reference, a LOOKUP model together with std images raises RuntimeError (it has no derivative to propagate them through).

    Integer ingest: a batch may carry the raw uint8 / uint16 camera codes (and a `datasets.StdSpec` instead of std images);
    CastTo + Normalize(max_val=code_max, default 255 / 65535) and the std synthesis then run on the device, and the stack
    crosses PCIe as 1-2 bytes per sample — bit-identical to handing over the CPU-transformed fp32 images.

    Like the reference, only the FIRST batch of the dataloader is measured (the `return` at :74 sits inside the
    loop), so pass the whole stack as one batch.
    """
    if not isinstance(dataloader, DataLoader):
        raise ArgumentTypeError(f"dataloader must be a torch DataLoader, got {type(dataloader)}")
    for flag in (use_uncertainty_weighting, use_relative_linearity_loss):
        if not isinstance(flag, bool):
            raise ArgumentTypeError("use_uncertainty_weighting / use_relative_linearity_loss must be bool")
    dev = as_device(device)
    table, interp_mode = model_table(icrf_model, dev)          # any InterpMode: the kernels evaluate all three
    for _, val_batch, std_batch, meta_batch in dataloader:
        images, stds = stage_batch(val_batch, std_batch, dev, code_max=code_max, expand_codes=True)
        exposures = meta_batch["exposure_time"]
        i_idx, j_idx, ratio_pairs = get_valid_exposure_pairs(exposures, RATIO_THRESHOLD)
        sums = kernels.pair_stats(images, stds, i_idx, j_idx, ratio_pairs, table, VALID_LO, VALID_HI,
                                  use_relative_linearity_loss, use_uncertainty_weighting, interp_mode=interp_mode)
        mean, std, errmean = spatial_statistics(sums, stds is not None)
        return ratio_pairs.to(dev), mean, std, errmean
    raise ValueError("the dataloader yielded no batches")

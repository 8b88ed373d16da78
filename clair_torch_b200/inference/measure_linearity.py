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
        exposures = meta_batch["exposure_time"]
        i_idx, j_idx, ratio_pairs = get_valid_exposure_pairs(exposures, RATIO_THRESHOLD)
        bands = _code_bands(val_batch, std_batch)
        if bands:
            sums, with_errors = _banded_code_statistics(val_batch, std_batch, bands, dev, code_max, i_idx, j_idx, ratio_pairs, table,
                                                        use_relative_linearity_loss, use_uncertainty_weighting, interp_mode)
        else:
            images, stds = stage_batch(val_batch, std_batch, dev, code_max=code_max, expand_codes=True)
            sums = kernels.pair_stats(images, stds, i_idx, j_idx, ratio_pairs, table, VALID_LO, VALID_HI,
                                      use_relative_linearity_loss, use_uncertainty_weighting, interp_mode=interp_mode)
            with_errors = stds is not None
        mean, std, errmean = spatial_statistics(sums, with_errors)
        return ratio_pairs.to(dev), mean, std, errmean
    raise ValueError("the dataloader yielded no batches")


_BAND_MIN_BYTES = 64 << 20      # below this the copy is too short to be worth cutting up


def _code_bands(val_batch, std_batch) -> int:
    """Number of row bands a page-locked uint8 / uint16 code batch (std: a StdSpec or none) is cut into so that its copy to the
    device overlaps the expansion + statistics of the bands already there; 0 = take the batch whole."""
    if not (torch.is_tensor(val_batch) and val_batch.dtype in (torch.uint8, torch.uint16) and val_batch.dim() == 4
            and not val_batch.is_cuda and val_batch.is_pinned() and val_batch.is_contiguous()):
        return 0
    if torch.is_tensor(std_batch) or val_batch.numel() * val_batch.element_size() < _BAND_MIN_BYTES:
        return 0
    h, w = val_batch.shape[-2:]
    for bands in (8, 6, 5, 4, 3, 2):
        if h % bands == 0 and ((h // bands) * w) % 4 == 0:
            return bands
    return 0


def _banded_code_statistics(codes, std_spec, bands, dev, code_max, i_idx, j_idx, ratio_pairs, table, relative, unc_weighting, interp_mode):
    """measure_linearity.py:41-72 for a page-locked code batch, pipelined: while band b is expanded (CastTo + Normalize + std
    synthesis) and its pair sums are accumulated, band b+1 crosses PCIe on a second stream (one strided copy per band,
    clair_copy_band_h2d; the stream is synchronised with the copies before the batch is let go).  The statistics kernel accumulates into
    the same (P, C, 5) sums with the band's table-row base, so the result is the whole-image one up to summation order."""
    n, c, h, w = codes.shape
    rows = h // bands
    main = torch.cuda.current_stream(dev)
    copier = torch.cuda.Stream(dev)
    staging = [torch.empty((n, c, rows, w), dtype=codes.dtype, device=dev) for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    sums = torch.zeros((len(i_idx), c, 5), dtype=torch.float64, device=dev)
    copier.wait_stream(main)
    with_errors = False
    for b in range(bands):
        k, r0 = b % 2, b * rows
        with torch.cuda.stream(copier):
            if b >= 2:
                copier.wait_event(consumed[k])
            kernels.copy_band_to_device(staging[k], codes, r0, copier)
            landed[k].record(copier)
        main.wait_event(landed[k])
        images, stds = kernels.expand_codes(staging[k], std_spec, code_max, dev)
        with_errors = stds is not None
        kernels.pair_stats(images, stds, i_idx, j_idx, ratio_pairs, table, VALID_LO, VALID_HI, relative, unc_weighting,
                           row_base=kernels.shard_row_base(c, h, w, r0), out=sums, interp_mode=interp_mode)
        consumed[k].record(main)
    for buf in staging:
        buf.record_stream(copier)
    # the copies read the pinned batch through a raw pointer: hold it until the last one has landed (the caller's loop variable
    # is the only other reference; the sums are read back by the caller right after anyway)
    landed[(bands - 1) % 2].synchronize()
    return sums, with_errors

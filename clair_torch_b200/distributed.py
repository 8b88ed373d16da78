"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in CPU tests).

The path shards naturally:
  * HDR merge / linearisation: by stack, or by row band of one huge stack — every output pixel depends only on its
    own N samples, so there is NO data-path collective;
  * linearity measurement / ICRF training: by row band; the only exchange is a sum-all-reduce of the (P, C, 5)
    float64 spatial sums and, when training, of the (C, L) float64 table gradient (a few KB, latency-bound).
    Per-band losses do NOT average to the full-image loss (sqrt of a sum of squared ratios of sums), which is why
    the sums are reduced rather than the losses.
Because the reference's LINEAR mode picks the table row from the flat NCHW index, every band has to
be told where it sits in the full frame: `band_row_base`.
"""
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import kernels


def stacks_for_rank(n_stacks: int, rank: int, world: int) -> list[int]:
    """Contiguous block partition of stack ids 0..n_stacks-1 (c4: 64 stacks over 8 GPUs -> 8 each)."""
    base, extra = divmod(n_stacks, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def row_band(height: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [r0, r1) of the full frame owned by `rank` (contiguous bands, remainder spread over the first ranks)."""
    base, extra = divmod(height, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def band_row_base(channels: int, full_height: int, width: int, first_row: int) -> np.ndarray:
    return kernels.shard_row_base(channels, full_height, width, first_row)


def take_band(stack: torch.Tensor, r0: int, r1: int) -> torch.Tensor:
    """Contiguous copy of rows [r0, r1) of an (N, C, H, W) stack."""
    return stack[:, :, r0:r1, :].contiguous()


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def train_icrf_step_data_parallel(icrf_model, optimizers, band_images, band_stds, exposures, row_base, group=None,
                                  **step_kwargs):
    """One training step on this rank's row band; sums and table gradient are all-reduced, so every rank applies the
    identical update to its replica of the (768-number) parameters.  Same keyword arguments as train_icrf_step."""
    from .training.icrf_training import train_icrf_step
    return train_icrf_step(icrf_model, optimizers, band_images, band_stds, exposures, row_base=row_base,
                           reduce_fn=lambda t: all_reduce_sum_(t, group), **step_kwargs)


def graphed_train_step_data_parallel(icrf_model, optimizers, band_images, band_stds, exposures, row_base, group=None,
                                     **step_kwargs):
    """train_icrf_step_data_parallel captured as one CUDA graph (kernels, both all-reduces, optimisers); call the returned
    object once per step.  Needs capturable optimisers and one eager step before (see GraphedTrainStep)."""
    from .training.icrf_training import GraphedTrainStep
    return GraphedTrainStep(icrf_model, optimizers, band_images, band_stds, exposures, row_base=row_base,
                            reduce_fn=lambda t: all_reduce_sum_(t, group), **step_kwargs)


def measure_linearity_band(band_images, band_stds, exposures, table: Optional[torch.Tensor], row_base,
                           use_uncertainty_weighting=True, use_relative_linearity_loss=True, group=None):
    """measure_linearity on a row band with the spatial sums all-reduced: every rank returns the full-image result."""
    from .common.general_functions import get_valid_exposure_pairs
    from .inference.measure_linearity import RATIO_THRESHOLD, VALID_HI, VALID_LO, spatial_statistics
    i_idx, j_idx, ratio = get_valid_exposure_pairs(exposures, RATIO_THRESHOLD)
    sums = kernels.pair_stats(band_images, band_stds, i_idx, j_idx, ratio, table, VALID_LO, VALID_HI,
                              use_relative_linearity_loss, use_uncertainty_weighting, row_base=row_base)
    all_reduce_sum_(sums, group)
    mean, std, err = spatial_statistics(sums, band_stds is not None)
    return ratio, mean, std, err

"""clair_torch_b200 — B200-native (sm_100a) implementation of the per-pixel radiometric hot path of
samivout/clair-torch: ICRF evaluation / linearisation, exposure-weighted HDR merge with first-order
uncertainty, and the pairwise exposure-ratio loss for ICRF training and linearity measurement.

The sub-packages mirror the reference's public surface for this path
(`clair_torch.models`, `clair_torch.inference`, `clair_torch.training`); the arithmetic runs in hand-written
CUDA kernels behind the C ABI of `include/clair_b200.h`.  There is no CPU or eager fallback.
"""
from . import _native, common, datasets, distributed, kernels, synthetic  # noqa: F401
from .common.enums import InterpMode
from .inference import (compute_hdr_image, compute_video_mean_and_std, linearize_dataset_generator,
                        measure_linearity)
from .models import ICRFModelBase, ICRFModelDirect, ICRFModelPCA
from .training import GraphedTrainStep, train_icrf, train_icrf_step

__version__ = "0.1.0"
__all__ = ["InterpMode", "ICRFModelBase", "ICRFModelDirect", "ICRFModelPCA", "compute_hdr_image", "linearize_dataset_generator",
           "measure_linearity", "compute_video_mean_and_std", "train_icrf", "train_icrf_step", "GraphedTrainStep"]

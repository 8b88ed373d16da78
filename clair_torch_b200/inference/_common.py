"""Shared plumbing of the driver loops: host->device staging of one collated batch and the model contract."""
from typing import Iterable, Optional

import torch

from ..common.enums import InterpMode
from ..models.base import ICRFModelBase


def as_device(device) -> torch.device:
    if not isinstance(device, (str, torch.device)):
        raise TypeError(f"device must be a str or torch.device, got {type(device)}")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"clair_torch_b200 runs on CUDA devices only (got device={device!r}); there is no CPU path")
    return dev


def stage_batch(val_batch, std_batch, device, gpu_transforms=None):
    """inference/hdr_merge.py:64-71: copy value / std images to the device and run the optional gpu transforms."""
    images = val_batch.to(device=device, non_blocking=True)
    stds = std_batch.to(device=device, non_blocking=True) if std_batch is not None else None
    for transform in gpu_transforms or ():
        if transform is not None:
            images = transform(images)
    return images, stds


def normalise_transforms(gpu_transforms) -> list:
    if gpu_transforms is None:
        return []
    if isinstance(gpu_transforms, Iterable):
        return list(gpu_transforms)
    return [gpu_transforms]


def linear_table(icrf_model: Optional[ICRFModelBase], device) -> Optional[torch.Tensor]:
    """The (C, L) table the fused kernels evaluate in LINEAR mode, or None for `icrf_model=None`."""
    if icrf_model is None:
        return None
    if not isinstance(icrf_model, ICRFModelBase):
        raise TypeError(f"icrf_model must be an ICRFModelBase, got {type(icrf_model)}")
    if icrf_model.interpolation_mode is not InterpMode.LINEAR:
        raise NotImplementedError("the fused B200 kernels evaluate the ICRF in InterpMode.LINEAR (the reference default); "
                                  f"got {icrf_model.interpolation_mode}")
    return icrf_model.icrf.detach().to(device=device, dtype=torch.float32)


def reject_artefacts(**datasets):
    for name, ds in datasets.items():
        if ds is not None:
            raise NotImplementedError(f"{name}: flat-field / dark-field corrections are the next row of the scope table "
                                      "(SURVEY.md §8(f) rank 1) and are not fused into the B200 kernels yet")

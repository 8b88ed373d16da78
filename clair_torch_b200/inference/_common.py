"""Shared plumbing of the driver loops: host->device staging of one collated batch and the model contract."""
from typing import Iterable, Optional

import torch

from ..common.enums import InterpMode
from ..models.base import ICRFModelBase
from ..common.errors import ArgumentTypeError


def as_device(device) -> torch.device:
    if not isinstance(device, (str, torch.device)):
        raise ArgumentTypeError(f"device must be a str or torch.device, got {type(device)}")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"clair_torch_b200 runs on CUDA devices only (got device={device!r}); there is no CPU path")
    return dev


def stage_batch(val_batch, std_batch, device, gpu_transforms=None, code_max=None, expand_codes=False):
    """inference/hdr_merge.py:64-71: copy value / std images to the device and run the optional gpu transforms.
    `expand_codes` — integer ingest for the pair drivers (the merge / linearise kernels ingest codes
    themselves): a uint8 / uint16 code batch is copied as it is and normalised on the device (`kernels.expand_codes`), with
    a `StdSpec` std batch synthesised there as well."""
    if expand_codes and torch.is_tensor(val_batch) and val_batch.dtype in (torch.uint8, torch.uint16):
        from .. import kernels
        if std_batch is not None and torch.is_tensor(std_batch):
            images, _ = kernels.expand_codes(val_batch, None, code_max, device)
            stds = std_batch.to(device=device, non_blocking=True)
        else:
            images, stds = kernels.expand_codes(val_batch, std_batch, code_max, device)
    else:
        if expand_codes and std_batch is not None and not torch.is_tensor(std_batch):
            raise ValueError("a StdSpec is evaluated on the device from integer codes: pass uint8 / uint16 value codes with it")
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


NATIVE_MODE = {InterpMode.LOOKUP: 1, InterpMode.LINEAR: 2, InterpMode.CATMULL: 3}     # CLAIR_INTERP_*


def model_table(icrf_model: Optional[ICRFModelBase], device):
    """(table (C, L) or None, CLAIR_INTERP_* code) of the model a driver was given (any interpolation mode)."""
    if icrf_model is None:
        return None, NATIVE_MODE[InterpMode.LINEAR]
    if not isinstance(icrf_model, ICRFModelBase):
        raise ArgumentTypeError(f"icrf_model must be an ICRFModelBase, got {type(icrf_model)}")
    return icrf_model.icrf.detach().to(device=device, dtype=torch.float32), NATIVE_MODE[icrf_model.interpolation_mode]


def check_artefact_dataset(name, ds):
    """Artefact datasets are duck-typed: the reference's Flat/DarkFieldArtefactMapDataset (file matching is host
    bookkeeping, out of scope) or datasets.InMemoryArtefactDataset — anything with get_matching_artefact_images."""
    if ds is not None and not hasattr(ds, "get_matching_artefact_images"):
        raise ArgumentTypeError(f"{name} must provide get_matching_artefact_images(frame_settings_list)")


def matching_dark_frames(main_dataset, dark_field_dataset, index_batch, device):
    """inference/hdr_merge.py:77-86: the dark frames matching the batch (None when nothing matches and the dataset
    is in skip mode)."""
    refs = [main_dataset.files[int(i)] for i in index_batch]
    _, dark_val, dark_std, _ = dark_field_dataset.get_matching_artefact_images(refs)
    if dark_val is None:
        return None, None
    return dark_val.to(device=device), (None if dark_std is None else dark_std.to(device=device))

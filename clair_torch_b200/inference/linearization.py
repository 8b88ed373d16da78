"""Single-image linearisation driver — call-compatible with clair_torch/inference/linearization.py:17-132."""
from typing import Generator

import torch
from torch.utils.data import DataLoader

from .. import kernels
from ..models.base import ICRFModelBase
from ._common import as_device, linear_table, normalise_transforms, reject_artefacts, stage_batch


def linearize_dataset_generator(dataloader: DataLoader, device, icrf_model: ICRFModelBase, flatfield_dataset=None,
                                gpu_transforms=None, dark_field_dataset=None
                                ) -> Generator[tuple[torch.Tensor, torch.Tensor, dict], None, None]:
    """Yields (linearised image, its uncertainty, metadata) per image, both on the CPU like the reference (:132).

    One kernel per image: f(x) and sigma = sqrt((f'(x) std)^2) are produced in the same pass (zeros when the
    dataset has no std images, :97).  The device->host copy uses pinned staging buffers.
    """
    if not isinstance(dataloader, DataLoader):
        raise TypeError(f"dataloader must be a torch DataLoader, got {type(dataloader)}")
    if not isinstance(icrf_model, ICRFModelBase):
        raise TypeError(f"icrf_model must be an ICRFModelBase, got {type(icrf_model)}")
    dev = as_device(device)
    if not dataloader.batch_size == 1:
        raise ValueError("For linearization only batch_size of 1 is allowed.")
    reject_artefacts(flatfield_dataset=flatfield_dataset, dark_field_dataset=dark_field_dataset)
    transforms = normalise_transforms(gpu_transforms)
    table = linear_table(icrf_model, dev)
    for _, val_batch, std_batch, meta_batch in dataloader:
        images, stds = stage_batch(val_batch, std_batch, dev, transforms)
        lin, sigma = kernels.linearize(images, stds, table)
        yield lin.squeeze().cpu(), sigma.squeeze().cpu(), meta_batch

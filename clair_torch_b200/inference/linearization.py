"""Single-image linearisation driver — call-compatible with clair_torch/inference/linearization.py:17-132."""
from typing import Generator

import torch
from torch.utils.data import DataLoader

from .. import kernels
from ..models.base import ICRFModelBase
from ._common import (as_device, check_artefact_dataset, matching_dark_frames, model_table, normalise_transforms,
                      stage_batch)
from ..common.errors import ArgumentTypeError


def linearize_dataset_generator(dataloader: DataLoader, device, icrf_model: ICRFModelBase, flatfield_dataset=None,
                                gpu_transforms=None, dark_field_dataset=None
                                ) -> Generator[tuple[torch.Tensor, torch.Tensor, dict], None, None]:
    """Yields (linearised image, its uncertainty, metadata) per image, both on the CPU like the reference (:132).

    One kernel per image: f(x) and sigma = sqrt((f'(x) std)^2) are produced in the same pass (zeros when the
    dataset has no std images, :97).  The device->host copy uses pinned staging buffers.
    """
    if not isinstance(dataloader, DataLoader):
        raise ArgumentTypeError(f"dataloader must be a torch DataLoader, got {type(dataloader)}")
    if not isinstance(icrf_model, ICRFModelBase):
        raise ArgumentTypeError(f"icrf_model must be an ICRFModelBase, got {type(icrf_model)}")
    dev = as_device(device)
    if not dataloader.batch_size == 1:
        raise ValueError("For linearization only batch_size of 1 is allowed.")
    check_artefact_dataset("flatfield_dataset", flatfield_dataset)
    check_artefact_dataset("dark_field_dataset", dark_field_dataset)
    main_dataset = dataloader.dataset
    transforms = normalise_transforms(gpu_transforms)
    table, interp_mode = model_table(icrf_model, dev)
    flat_val = flat_std = None
    if flatfield_dataset is not None:                     # linearization.py:50-57: one flat field for the whole run
        _, flat_val, flat_std, _ = flatfield_dataset.get_matching_artefact_images([main_dataset.files[0]])
    for index_batch, val_batch, std_batch, meta_batch in dataloader:
        plain = dark_field_dataset is None and flat_val is None and not transforms
        codes = val_batch.dtype in (torch.uint8, torch.uint16)
        if codes and not plain:
            raise NotImplementedError("integer-code batches are linearised without artefact corrections / gpu transforms: "
                                      "hand over normalised fp32 images for those")
        std_is_tensor = torch.is_tensor(std_batch)
        if std_batch is not None and not std_is_tensor and not codes:
            raise ValueError("a StdSpec is evaluated by the integer-ingest kernel: pass uint8 / uint16 value codes with it")
        zero_copy = (plain and not val_batch.is_cuda and val_batch.is_pinned() and val_batch.is_contiguous()
                     and (not std_is_tensor or (std_batch.is_pinned() and std_batch.is_contiguous())))
        if zero_copy:
            images, stds = val_batch, std_batch          # read over PCIe by the kernel itself
        elif codes:                                      # integer ingest: 1-2 bytes per sample cross PCIe
            images = val_batch.to(device=dev, non_blocking=True)
            stds = std_batch.to(device=dev, non_blocking=True) if std_is_tensor else std_batch
        else:
            images, stds = stage_batch(val_batch, std_batch, dev, transforms)
        if dark_field_dataset is not None:                # linearization.py:73-91,108-116
            dark_val, dark_std = matching_dark_frames(main_dataset, dark_field_dataset, index_batch, dev)
            if dark_val is not None:
                images, stds = kernels.dark_field_mix(images, stds, dark_val, dark_std)
        if plain:
            # results go straight to page-locked host memory (the reference ends with .cpu(), :132)
            lin, sigma = kernels.linearize(images, stds, table, device=dev, pinned_out=True, interp_mode=interp_mode)
            torch.cuda.current_stream(dev).synchronize()
            yield lin.squeeze(), sigma.squeeze(), meta_batch
            continue
        lin, sigma = kernels.linearize(images, stds, table, interp_mode=interp_mode)
        if flat_val is not None:                          # linearization.py:118-130: the mean is a constant here
            kernels.flat_field_correct_(lin, sigma, flat_val, flat_std, mean_in_graph=False)
        yield lin.squeeze().cpu(), sigma.squeeze().cpu(), meta_batch

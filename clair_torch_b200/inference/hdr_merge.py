"""HDR merge driver — call-compatible with clair_torch/inference/hdr_merge.py:19-155."""
import collections
from typing import Callable, Optional

import torch
from torch.utils.data import DataLoader

from .. import kernels
from ..models.base import ICRFModelBase
from ._common import (as_device, check_artefact_dataset, matching_dark_frames, model_table, normalise_transforms,
                      stage_batch)
from ..common.errors import ArgumentTypeError


def compute_hdr_image(dataloader: DataLoader, device, icrf_model: Optional[ICRFModelBase] = None,
                      weight_fn: Optional[Callable] = None, flat_field_dataset=None, gpu_transforms=None,
                      dark_field_dataset=None, *, radiance_dtype: Optional[torch.dtype] = None, host_out=None,
                      code_max: Optional[float] = None, staged: Optional[bool] = None, code_layout: str = "planar"):
    """Exposure-weighted HDR merge of a stationary exposure stack with first-order uncertainty.

    Each DataLoader batch goes through ONE fused kernel (ICRF evaluation, Gaussian weights, weighted running mean
    and the closed-form variance of the reference's autograd pass).  As in the reference, `weight_fn` is only
    tested against None: any callable selects `gaussian_value_weights` with scale 30 (hdr_merge.py:95).

    Returns (radiance (C,H,W) squeezed, sigma (C,H,W) squeezed or None when the batches carry no std images).
    `radiance_dtype` defaults to the dtype the reference returns (float64, because the collated exposure times
    are float64 — SURVEY.md Q6); pass torch.float32 to halve the output traffic.

    Host batches that are page-locked (DataLoader(pin_memory=True) or pre-pinned tensors) are streamed band by band:
    the copy engine moves the next band of every frame to the device while the kernel merges the current one
    (`staged=False`: the kernel reads the host memory itself over PCIe instead); pageable batches are copied to the
    device first like the reference does.  `host_out=(radiance, sigma)`, two pinned (C,H,W) host tensors, makes the
    kernel write the results straight to host memory (they are then what is returned, and the call synchronises the
    stream before returning them, as it does whenever it read a host-resident batch in place).

    Models in any InterpMode are accepted: LINEAR (the reference default) runs the fused fast kernels, LOOKUP and
    CATMULL an all-modes kernel.  As in the reference, a LOOKUP model has no derivative with respect to the image, so
    with std images its uncertainty comes from the weights alone, and without `weight_fn` there is nothing to
    differentiate: the reference's autograd call raises there and so does this function (RuntimeError).

    Integer ingest: batches may carry the raw uint8 / uint16 camera codes instead of
    normalised fp32 images; the kernel then performs the reference's CastTo(float32) + Normalize(max_val=code_max,
    min_val=0) on load (code_max defaults to 255 / 65535), and the std batch may be a `datasets.StdSpec`
    (MissingStdMode.MULTIPLIER / CONSTANT evaluated in-register).  Results are bit-identical to feeding the
    CPU-transformed fp32 images; 4-8x fewer bytes cross PCIe and HBM.  With `code_layout="hwc_bgr"` the code batches are
    (N, H, W, 3) BGR — `cv2.imread` output, stacked — and CvToTorch is fused into the load as well.
    """
    if not isinstance(dataloader, DataLoader):
        raise ArgumentTypeError(f"dataloader must be a torch DataLoader, got {type(dataloader)}")
    if weight_fn is not None and not callable(weight_fn):
        raise ArgumentTypeError("weight_fn must be callable or None")
    dev = as_device(device)
    check_artefact_dataset("flat_field_dataset", flat_field_dataset)
    check_artefact_dataset("dark_field_dataset", dark_field_dataset)
    main_dataset = dataloader.dataset
    transforms = normalise_transforms(gpu_transforms)
    table, interp_mode = model_table(icrf_model, dev)

    state = kernels.HdrMergeState()
    result = None
    # Page-locked host batches are read by the kernels / the copy engine through raw pointers, which torch's caching host
    # allocator knows nothing about: every such batch is kept alive here until an event recorded behind its last reader
    # has completed (a DataLoader pin thread would otherwise get the block back and overwrite it mid-read).
    in_flight = collections.deque()
    batches = iter(dataloader)
    current = next(batches, None)
    while current is not None:
        upcoming = next(batches, None)                   # look one batch ahead to know which one is the last
        _, val_batch, std_batch, meta_batch = current
        std_is_tensor = torch.is_tensor(std_batch)
        index_batch = current[0]
        zero_copy = (not transforms and dark_field_dataset is None and not val_batch.is_cuda and val_batch.is_pinned()
                     and val_batch.is_contiguous()
                     and (not std_is_tensor or (std_batch.is_pinned() and std_batch.is_contiguous())))
        if zero_copy:
            images, stds = val_batch, std_batch
        else:
            images, stds = stage_batch(val_batch, std_batch if std_is_tensor else None, dev, transforms)
            stds = stds if std_is_tensor else std_batch
        fused_dark = None
        if dark_field_dataset is not None:
            # hot pixels of the matching dark frames select a blurred copy; image-std and dark-std variance terms fold
            # into one effective std (hdr_merge.py:76-92,107-126)
            if images.dtype != torch.float32:
                raise NotImplementedError("dark-field correction takes normalised fp32 images (not raw integer codes)")
            dark_val, dark_std = matching_dark_frames(main_dataset, dark_field_dataset, index_batch, dev)
            if dark_val is not None:
                if kernels.can_fuse_dark(images, stds, dark_std, interp_mode if table is not None else 2):
                    fused_dark = (dark_val, dark_std)          # mixed in registers by the merge kernel itself
                else:
                    images, stds = kernels.dark_field_mix(images, stds if torch.is_tensor(stds) else None, dark_val, dark_std)
        if table is not None and interp_mode == 1 and stds is not None and weight_fn is None:
            # hdr_merge.py:107-112: autograd.grad of a mean that does not depend on the images
            raise RuntimeError("a LOOKUP model without weight_fn leaves the merged image independent of the input images: "
                               "there is no gradient to propagate the std images through (the reference raises here too)")
        exposures = meta_batch["exposure_time"]
        out_dtype = radiance_dtype
        if out_dtype is None:
            value_dtype = torch.float32 if not images.dtype.is_floating_point else images.dtype
            out_dtype = torch.promote_types(value_dtype, exposures.dtype if torch.is_tensor(exposures) else torch.float64)
        result = kernels.hdr_merge_update(state, images, stds, exposures, table, weight_fn is not None,
                                          is_final=upcoming is None, radiance_dtype=out_dtype, device=dev,
                                          host_out=host_out if (upcoming is None and flat_field_dataset is None) else None,
                                          code_max=code_max, interp_mode=interp_mode,
                                          staged=staged if not images.is_cuda else None, dark=fused_dark,
                                          code_layout=code_layout)
        if not images.is_cuda:
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(dev))
            in_flight.append((done, images, stds))
            while in_flight and in_flight[0][0].query():
                in_flight.popleft()
        current = upcoming
    if result is None:
        raise ValueError("the dataloader yielded no batches")
    radiance, sigma = result
    if flat_field_dataset is not None:
        # hdr_merge.py:131-153: the whole-image mean of the flat field is inside the autograd graph there
        _, flat_val, flat_std, _ = flat_field_dataset.get_matching_artefact_images([main_dataset.files[0]])
        kernels.flat_field_correct_(radiance, sigma, flat_val, flat_std if sigma is not None else None, mean_in_graph=True)
        if host_out is not None:
            host_out[0].copy_(radiance, non_blocking=True)
            if sigma is not None:
                host_out[1].copy_(sigma, non_blocking=True)
            radiance, sigma = host_out[0], (host_out[1] if sigma is not None else None)
    if in_flight or not radiance.is_cuda:
        # host-resident inputs must outlive their readers, and host-resident results are returned ready to use (the
        # reference's results are; INTEGRATION.md "synchronisation")
        torch.cuda.current_stream(dev).synchronize()
        in_flight.clear()
    return radiance.squeeze(), (sigma.squeeze() if sigma is not None else None)

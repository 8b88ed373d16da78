"""Mean and standard error of video frames — call-compatible with
clair_torch/inference/inferential_statistics.py:19-49."""
import math
from typing import Optional

import torch
from torch.utils.data import DataLoader

from ..common.enums import VarianceMode
from ..common.statistics import WBOMeanVar
from ..models.base import ICRFModelBase
from ._common import as_device, model_table
from ..common.errors import ArgumentTypeError


def compute_video_mean_and_std(dataloader: DataLoader, device, icrf_model: Optional[ICRFModelBase] = None):
    """(mean (C,H,W), std of the mean (C,H,W)) over all frames of all batches; frames are optionally linearised
    first.  One fused pass per batch (ICRF + running mean / M2 merge)."""
    if not isinstance(dataloader, DataLoader):
        raise ArgumentTypeError(f"dataloader must be a torch DataLoader, got {type(dataloader)}")
    dev = as_device(device)
    table, interp_mode = model_table(icrf_model, dev)
    handler = WBOMeanVar(dim=0, variance_mode=VarianceMode.SAMPLE_FREQUENCY)
    number_of_frames = 0
    for _, val_batch, _, _ in dataloader:
        frames = val_batch.to(device=dev, non_blocking=True)
        number_of_frames += frames.shape[0]
        handler.update_values(frames, None, table=table, interp_mode=interp_mode)
    if number_of_frames == 0:
        raise ValueError("the dataloader yielded no batches")
    return handler.mean.squeeze(), torch.sqrt(handler.variance().squeeze()) / math.sqrt(number_of_frames)

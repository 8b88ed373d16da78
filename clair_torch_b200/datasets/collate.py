"""Batch format consumed by the driver loops — same tuple as clair_torch/datasets/collate.py:8-43."""
import torch
from torch.utils.data._utils.collate import default_collate

from .stack_dataset import StdSpec


def custom_collate(batch):
    """(index_batch int64 (N,), val_batch (N,C,H,W), std_batch (N,C,H,W) | None, meta_batch dict of (N,) tensors).

    The batch is sorted by ascending exposure time; python-float exposure times collate to float64
    (SURVEY.md Q6); one missing std image makes the whole std batch None.
    """
    ordered = sorted(batch, key=lambda item: item[3]["exposure_time"])
    indices, vals, stds, metas = zip(*ordered)
    if any(s is None for s in stds):
        std_batch = None
    elif all(isinstance(s, StdSpec) for s in stds):
        if any(s != stds[0] for s in stds):
            raise ValueError("all frames of a batch must share one StdSpec")
        std_batch = stds[0]                       # evaluated in the kernel, nothing to stack
    else:
        std_batch = default_collate(stds)
    return default_collate(indices), default_collate(vals), std_batch, default_collate(metas)

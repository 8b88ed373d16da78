"""In-memory map-style dataset yielding the reference's item tuple.

The reference's file-backed datasets (clair_torch/datasets/base.py:20-172) are out of scope; anything that
yields `(index, val (C,H,W) fp32, std (C,H,W) fp32 | None, {'exposure_time': float})` works with the driver
loops here, including the reference's own ImageMapDataset when clair_torch is installed.
"""
from typing import Optional, Sequence

import torch
from torch.utils.data import Dataset


class ExposureStackDataset(Dataset):
    def __init__(self, vals: Sequence[torch.Tensor] | torch.Tensor, stds: Optional[Sequence[torch.Tensor] | torch.Tensor],
                 exposures: Sequence[float], copy: bool = False):
        if len(vals) != len(exposures) or (stds is not None and len(stds) != len(vals)):
            raise ValueError("vals, stds and exposures must have the same length")
        self.vals, self.stds, self.exposures, self.copy = vals, stds, [float(e) for e in exposures], copy
        self.files = tuple(range(len(vals)))

    def __len__(self) -> int:
        return len(self.vals)

    def __getitem__(self, idx: int):
        val = self.vals[idx]
        std = None if self.stds is None else self.stds[idx]
        if self.copy:
            val = val.clone()
            std = None if std is None else std.clone()
        return idx, val, std, {"exposure_time": self.exposures[idx]}

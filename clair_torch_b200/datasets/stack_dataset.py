"""In-memory map-style dataset yielding the reference's item tuple.

The reference's file-backed datasets (clair_torch/datasets/base.py:20-172) are out of scope; anything that
yields `(index, val (C,H,W) fp32, std (C,H,W) fp32 | None, {'exposure_time': float})` works with the driver
loops here, including the reference's own ImageMapDataset when clair_torch is installed.
"""
from dataclasses import dataclass
from typing import Optional, Sequence

import torch
from torch.utils.data import Dataset


@dataclass(frozen=True)
class StdSpec:
    """A std image that is a function of the value image, evaluated inside the kernel instead of being materialised:
    the two ways MultiFileMapDataset fills in a missing std image (clair_torch/datasets/base.py:128-133).
      mode "multiplier": std = value * `value`     (MissingStdMode.MULTIPLIER; value after normalisation)
      mode "constant"  : std = `value`             (MissingStdMode.CONSTANT)
    """
    mode: str
    value: float

    def __post_init__(self):
        if self.mode not in ("multiplier", "constant"):
            raise ValueError(f"StdSpec mode must be 'multiplier' or 'constant', got {self.mode!r}")


class ExposureStackDataset(Dataset):
    def __init__(self, vals: Sequence[torch.Tensor] | torch.Tensor,
                 stds: Optional[Sequence[torch.Tensor] | torch.Tensor | StdSpec], exposures: Sequence[float],
                 copy: bool = False):
        """`vals` may be fp32 (already normalised, the reference's hand-over) or raw uint8 / uint16 codes (integer
        ingest); `stds` a matching sequence of fp32 images, None, or a StdSpec."""
        if len(vals) != len(exposures) or (stds is not None and not isinstance(stds, StdSpec) and len(stds) != len(vals)):
            raise ValueError("vals, stds and exposures must have the same length")
        self.vals, self.stds, self.exposures, self.copy = vals, stds, [float(e) for e in exposures], copy
        self.files = tuple(range(len(vals)))

    def __len__(self) -> int:
        return len(self.vals)

    def __getitem__(self, idx: int):
        val = self.vals[idx]
        if self.stds is None or isinstance(self.stds, StdSpec):
            std = self.stds
        else:
            std = self.stds[idx]
        if self.copy:
            val = val.clone()
            std = std.clone() if torch.is_tensor(std) else std
        return idx, val, std, {"exposure_time": self.exposures[idx]}


class InMemoryArtefactDataset:
    """Flat-field / dark-field calibration frames held in memory.  The reference matches artefact files to frames by
    filename metadata (clair_torch/datasets/base.py:175-296, host bookkeeping that is out of scope); here frame `i` of
    the main dataset matches artefact `i mod len`, and the return value has the reference's collated structure."""

    def __init__(self, vals: Sequence[torch.Tensor], stds: Optional[Sequence[torch.Tensor]], exposures: Optional[Sequence[float]] = None):
        self.vals, self.stds = vals, stds
        self.exposures = [float(e) for e in exposures] if exposures is not None else None

    def __len__(self) -> int:
        return len(self.vals)

    def get_matching_artefact_images(self, reference_frame_settings_list):
        from .collate import custom_collate
        items = []
        for pos, ref in enumerate(reference_frame_settings_list):
            i = int(ref) % len(self.vals)
            # custom_collate sorts by exposure time; without exposures the request order is kept
            key = self.exposures[i] if self.exposures is not None else float(pos)
            items.append((i, self.vals[i], None if self.stds is None else self.stds[i], {"exposure_time": key}))
        return custom_collate(items)

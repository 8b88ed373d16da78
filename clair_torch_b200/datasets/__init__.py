from .collate import custom_collate
from .stack_dataset import ExposureStackDataset, StdSpec

__all__ = ["custom_collate", "ExposureStackDataset", "StdSpec"]

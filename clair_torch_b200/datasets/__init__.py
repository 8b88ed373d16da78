from .collate import custom_collate
from .stack_dataset import ExposureStackDataset

__all__ = ["custom_collate", "ExposureStackDataset"]

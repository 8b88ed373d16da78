from .collate import custom_collate
from .stack_dataset import ExposureStackDataset, InMemoryArtefactDataset, StdSpec

__all__ = ["custom_collate", "ExposureStackDataset", "InMemoryArtefactDataset", "StdSpec"]

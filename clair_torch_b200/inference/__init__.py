from .hdr_merge import compute_hdr_image
from .linearization import linearize_dataset_generator
from .measure_linearity import measure_linearity

__all__ = ["compute_hdr_image", "linearize_dataset_generator", "measure_linearity"]

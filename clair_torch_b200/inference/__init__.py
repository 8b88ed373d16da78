from .hdr_merge import compute_hdr_image
from .inferential_statistics import compute_video_mean_and_std
from .linearization import linearize_dataset_generator
from .measure_linearity import measure_linearity

__all__ = ["compute_hdr_image", "compute_video_mean_and_std", "linearize_dataset_generator", "measure_linearity"]

from .enums import InterpMode, VarianceMode
from .general_functions import (get_pairwise_valid_pixel_mask, get_valid_exposure_pairs, weighted_mean_and_std)

__all__ = ["InterpMode", "VarianceMode", "get_valid_exposure_pairs", "get_pairwise_valid_pixel_mask", "weighted_mean_and_std"]

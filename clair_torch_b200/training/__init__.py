from .icrf_training import GraphedTrainStep, linearity_loss_and_table_grad, train_icrf, train_icrf_step
from .losses import (combined_gaussian_pair_weights, compute_endpoint_penalty, compute_monotonicity_penalty,
                     compute_range_penalty, compute_smoothness_penalty, compute_spatial_linearity_loss,
                     gaussian_value_weights, pixelwise_linearity_loss)

__all__ = ["train_icrf", "train_icrf_step", "GraphedTrainStep", "linearity_loss_and_table_grad", "gaussian_value_weights",
           "combined_gaussian_pair_weights", "pixelwise_linearity_loss", "compute_spatial_linearity_loss",
           "compute_monotonicity_penalty", "compute_smoothness_penalty", "compute_range_penalty",
           "compute_endpoint_penalty"]

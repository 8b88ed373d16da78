from .base import ICRFModelBase
from .icrf_model import ICRFModelDirect, ICRFModelPCA

__all__ = ["ICRFModelBase", "ICRFModelDirect", "ICRFModelPCA"]

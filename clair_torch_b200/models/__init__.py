from .base import ICRFModelBase
from .icrf_model import ICRFModelDirect

__all__ = ["ICRFModelBase", "ICRFModelDirect"]

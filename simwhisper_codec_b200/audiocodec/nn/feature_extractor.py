"""Import-path mirror of the reference's `audiocodec/nn/feature_extractor.py` (`audiocodec/model.py:10`)."""
from ..model import MelFeatureExtractor  # noqa: F401

__all__ = ["MelFeatureExtractor"]

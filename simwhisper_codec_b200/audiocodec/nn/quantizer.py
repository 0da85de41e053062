"""Import-path mirror of the reference's `audiocodec/nn/quantizer.py` (`audiocodec/model.py:12`)."""
from ..model import GroupFiniteScalarQuantizer  # noqa: F401

__all__ = ["GroupFiniteScalarQuantizer"]

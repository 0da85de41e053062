"""B200-native SimWhisper-Codec hot path (sm_100a CUDA kernels behind the reference's AudioCodec API)."""
from .audiocodec.model import AudioCodec  # noqa: F401

__all__ = ["AudioCodec"]

"""Import-path mirror of the reference's `audiocodec/nn/modules.py` (the classes `audiocodec/model.py:11` imports):
`from audiocodec.nn.modules import OmniAudioEncoder, OmniAudioDecoder, FrameStackDownConv, FrameStackUpConv, Vocos`.
The classes are the C-ABI shells defined in ..model; the reference's other (dead-code) modules have no kernels."""
from ..model import FrameStackDownConv, FrameStackUpConv, OmniAudioDecoder, OmniAudioEncoder, Vocos  # noqa: F401

__all__ = ["OmniAudioEncoder", "OmniAudioDecoder", "FrameStackDownConv", "FrameStackUpConv", "Vocos"]

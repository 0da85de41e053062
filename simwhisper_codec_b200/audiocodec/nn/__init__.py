"""Mirror of the reference's audiocodec/nn package: the module classes live in ..model and are thin
shells over the C ABI (include/swc.h)."""
from ..model import (FrameStackDownConv, FrameStackUpConv, GroupFiniteScalarQuantizer, MelFeatureExtractor,  # noqa: F401
                     OmniAudioDecoder, OmniAudioEncoder, Vocos)

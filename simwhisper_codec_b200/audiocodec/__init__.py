from .model import AudioCodec  # noqa: F401

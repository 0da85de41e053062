"""Host-side helpers mirroring the reference's utils/helpers.py for the inference CLI (audio I/O, logging)."""
from .helpers import find_audio_files, load_audio, save_audio, set_logging  # noqa: F401

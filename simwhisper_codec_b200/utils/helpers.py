"""Audio I/O for the inference CLI — the counterpart of the reference's utils/helpers.py:60-111
(`set_logging`, `load_audio`, `save_audio`, `find_audio_files`) without torchaudio's file backends
(`torchaudio.load` needs torchcodec, which this image does not have).

* WAV container parsing / writing is done here (RIFF/WAVE PCM 8/16/24/32-bit and IEEE float 32/64, any channel count).
* `load_audio` averages channels to mono and resamples to the codec rate with the same windowed-sinc polyphase filter as
  `torchaudio.functional.resample` defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99), which is what the
  reference calls (helpers.py:86-87).  The convolution runs as one torch conv1d on whichever device is asked for; this is
  pre-processing outside the codec hot path.
* `save_audio` writes 16-bit signed PCM like the reference (`encoding='PCM_S', bits_per_sample=16`, helpers.py:95-102).
"""
from __future__ import annotations

import glob
import logging
import math
import os
import struct
import sys
from typing import List, Tuple

import numpy as np
import torch


def set_logging(level="INFO"):
    """reference utils/helpers.py:60-75."""
    if isinstance(level, str):
        level = getattr(logging, level.upper(), logging.INFO)
    rank = os.environ.get("RANK", 0)
    logging.basicConfig(level=level, stream=sys.stdout,
                        format=f"%(asctime)s [RANK {rank}] (%(module)s:%(lineno)d) %(levelname)s : %(message)s")


# --------------------------------------------------------------------------------------------- WAV container
def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """Returns (float32 array of shape (channels, frames) in [-1, 1), sample_rate)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None or len(fmt) < 16:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == 0xFFFE and len(fmt) >= 26:        # WAVE_FORMAT_EXTENSIBLE: the real tag is the first 2 bytes of the sub-format GUID
        tag = struct.unpack("<H", fmt[24:26])[0]
    if ch < 1:
        raise ValueError(f"{path}: bad channel count {ch}")
    bps = bits // 8
    n = len(payload) // (bps * ch)
    raw = payload[: n * bps * ch]
    if tag == 1:        # integer PCM
        if bits == 8:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = v.astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:      # IEEE float
        if bits == 32:
            x = np.frombuffer(raw, dtype="<f4").astype(np.float32)
        elif bits == 64:
            x = np.frombuffer(raw, dtype="<f8").astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported float width {bits}")
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag} (only PCM and IEEE float)")
    return np.ascontiguousarray(x.reshape(n, ch).T), int(rate)


def pcm16_from_float(x: np.ndarray) -> np.ndarray:
    """float [-1, 1) -> int16 with rounding and clipping (sox/torchaudio PCM_S 16 conversion)."""
    return np.clip(np.rint(np.asarray(x, dtype=np.float64) * 32768.0), -32768, 32767).astype("<i2")


def write_wav_pcm16(path: str, audio: np.ndarray, sample_rate: int) -> None:
    """audio: (channels, frames) or (frames,) float."""
    a = np.asarray(audio)
    if a.ndim == 1:
        a = a[None, :]
    ch, n = a.shape
    pcm = pcm16_from_float(a.T).tobytes()              # interleaved
    hdr = b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, sample_rate, sample_rate * ch * 2, ch * 2, 16) + b"data" + struct.pack("<I", len(pcm))
    with open(path, "wb") as f:
        f.write(hdr + pcm)


# --------------------------------------------------------------------------------------------- resampling
def sinc_resample_kernel(orig: int, new: int, lowpass_filter_width: int = 6, rolloff: float = 0.99,
                         dtype=torch.float32, device=None):
    """Polyphase windowed-sinc (Hann) kernel, as torchaudio.functional.resample builds it.
    Returns (kernels (new, 1, 2*width + orig), width, orig, new) with orig/new reduced by their gcd."""
    g = math.gcd(int(orig), int(new))
    orig, new = int(orig) // g, int(new) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64, device=device)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64, device=device)[:, None, None] / new + idx
    t = (t * base_freq).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.ones_like(t), t.sin() / t) * window * scale
    return kernels.to(dtype), width, orig, new


def resample(wav: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """wav (..., time) -> (..., ceil(new * time / orig)); same result as torchaudio.functional.resample defaults."""
    if orig_freq == new_freq:
        return wav
    kernels, width, orig, new = sinc_resample_kernel(orig_freq, new_freq, dtype=wav.dtype, device=wav.device)
    shape = wav.shape
    x = wav.reshape(-1, shape[-1])
    length = x.shape[-1]
    x = torch.nn.functional.pad(x, (width, width + orig))
    y = torch.nn.functional.conv1d(x[:, None], kernels, stride=orig)           # (n, new, frames)
    y = y.transpose(1, 2).reshape(x.shape[0], -1)
    target = math.ceil(new * length / orig)
    return y[..., :target].reshape(shape[:-1] + (target,))


# --------------------------------------------------------------------------------------------- reference API
def load_audio(audio_path: str, target_sample_rate: int, device="cpu") -> torch.Tensor:
    """reference utils/helpers.py:77-93 — mono mix, resample, shape (1, 1, time) float32."""
    ext = os.path.splitext(audio_path)[1].lower()
    if ext == ".wav":
        x, rate = read_wav(audio_path)
        wav = torch.from_numpy(x).to(device)
    else:
        wav, rate = _decode_with_external_backend(audio_path)
        wav = wav.to(device)
    if wav.shape[0] > 1:
        wav = wav.mean(dim=0, keepdim=True)
    if rate != target_sample_rate:
        wav = resample(wav, rate, target_sample_rate)
    return wav.reshape(1, 1, -1).contiguous()


def _decode_with_external_backend(audio_path: str):
    """flac / mp3 (the reference reads them through torchaudio.load, utils/helpers.py:83): use soundfile or torchaudio when
    one of them works in this environment, else raise (the built-in reader only knows RIFF/WAVE)."""
    try:
        import soundfile as sf
        data, rate = sf.read(audio_path, dtype="float32", always_2d=True)
        return torch.from_numpy(data.T.copy()), int(rate)
    except Exception:
        pass
    try:
        import torchaudio
        wav, rate = torchaudio.load(audio_path)
        return wav.to(torch.float32), int(rate)
    except Exception as e:
        raise RuntimeError(f"{audio_path}: only RIFF/WAVE input is supported by the built-in reader and no external decoder "
                           f"(soundfile / torchaudio) is usable here: {str(e)[:120]}")


def save_audio(audio_outpath: str, audio_out, sample_rate: int) -> None:
    """reference utils/helpers.py:95-103 — 16-bit signed PCM WAV; audio_out (channels, time) tensor/array."""
    a = audio_out.detach().cpu().float().numpy() if isinstance(audio_out, torch.Tensor) else np.asarray(audio_out)
    write_wav_pcm16(audio_outpath, a, int(sample_rate))
    logging.info(f"Successfully saved audio at {audio_outpath}")


def find_audio_files(input_dir: str) -> List[str]:
    """reference utils/helpers.py:105-111."""
    out: List[str] = []
    for ext in ("*.flac", "*.mp3", "*.wav"):
        out.extend(glob.glob(os.path.join(input_dir, "**", ext), recursive=True))
    logging.info(f"Found {len(out)} audio files in {input_dir}")
    return sorted(out)


def can_decode(audio_path: str) -> bool:
    """True for RIFF/WAVE, and for flac / mp3 when an external decoder is importable; the CLI skips (with a warning) what
    cannot be decoded instead of aborting the whole directory on the first such file."""
    if os.path.splitext(audio_path)[1].lower() == ".wav":
        return True
    for mod in ("soundfile", "torchaudio"):
        try:
            __import__(mod)
            return True
        except Exception:
            continue
    return False

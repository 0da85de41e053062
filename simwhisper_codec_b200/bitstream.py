"""Wire format for Group-FSQ codes.

The reference keeps codes as int32 tensors of shape (8, T) with values in [0, 2016) (quantizer.py:169-179: 8 groups,
levels [8,7,6,6] -> 2016 = 8*7*6*6 codes per group).  2016 <= 2^11, so one 80 ms frame is 8 x 11 = 88 bits = 11 bytes:
137.5 B/s = 1.1 kbit/s, the bitrate the reference quotes (README.md:25).

Stream layout (little-endian):
    magic  b"SWC1"            4 bytes
    groups u8, bits u8        2 bytes   (8, 11)
    codebook_size u16         2 bytes   (2016)
    frames u32                4 bytes
    payload                   ceil(frames * groups * bits / 8) bytes; frame-major, group-minor, each code `bits` wide,
                              packed LSB-first into a little-endian bit stream
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"SWC1"
HEADER = struct.Struct("<4sBBHI")


def pack_codes(codes, codebook_size: int = 2016) -> bytes:
    """codes: (groups, frames) integer array/tensor -> bytes."""
    c = np.asarray(codes.cpu() if hasattr(codes, "cpu") else codes)
    if c.ndim != 2:
        raise ValueError(f"expected (groups, frames), got shape {c.shape}")
    groups, frames = c.shape
    bits = max(1, int(codebook_size - 1).bit_length())
    if groups > 255:
        raise ValueError("too many groups")
    if c.size and (int(c.min()) < 0 or int(c.max()) >= codebook_size):
        raise ValueError(f"code out of range [0, {codebook_size})")
    flat = c.T.astype(np.uint32).reshape(-1)                                    # frame-major, group-minor
    bitmat = ((flat[:, None] >> np.arange(bits, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)
    payload = np.packbits(bitmat.reshape(-1), bitorder="little").tobytes()
    return HEADER.pack(MAGIC, groups, bits, codebook_size, frames) + payload


def unpack_codes(blob: bytes) -> np.ndarray:
    """bytes -> (groups, frames) int32."""
    if len(blob) < HEADER.size:
        raise ValueError("truncated stream")
    magic, groups, bits, codebook_size, frames = HEADER.unpack_from(blob, 0)
    if magic != MAGIC:
        raise ValueError("bad magic")
    n = frames * groups
    need = (n * bits + 7) // 8
    payload = np.frombuffer(blob, dtype=np.uint8, count=need, offset=HEADER.size) if need else np.zeros(0, np.uint8)
    if payload.size < need:
        raise ValueError("truncated payload")
    bitvec = np.unpackbits(payload, bitorder="little")[: n * bits].reshape(n, bits).astype(np.uint32)
    vals = (bitvec << np.arange(bits, dtype=np.uint32)[None, :]).sum(axis=1)
    if n and int(vals.max()) >= codebook_size:
        raise ValueError("code out of range in stream")
    return vals.reshape(frames, groups).T.astype(np.int32).copy()


def bitrate_bps(groups: int = 8, codebook_size: int = 2016, frame_rate_hz: float = 12.5) -> float:
    return groups * max(1, int(codebook_size - 1).bit_length()) * frame_rate_hz

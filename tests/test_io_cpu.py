"""CPU tests of the host-side rows next to the hot path: WAV I/O + resampling of the
inference CLI and the 11-bit code wire format."""
import math

import numpy as np
import pytest
import torch

from simwhisper_codec_b200 import bitstream
from simwhisper_codec_b200.utils import helpers


def test_wav_roundtrip_pcm16(tmp_path):
    g = np.random.default_rng(0)
    x = np.clip(g.normal(0, 0.3, size=(2, 5000)), -1, 1).astype(np.float32)
    p = str(tmp_path / "a.wav")
    helpers.write_wav_pcm16(p, x, 24000)
    y, rate = helpers.read_wav(p)
    assert rate == 24000 and y.shape == x.shape
    assert np.abs(y - x).max() <= 1.0 / 32768 + 1e-7            # half an LSB of rounding, one LSB at the +1.0 clip point
    # values representable in 16 bits survive exactly, and the clip points behave like sox/torchaudio PCM_S 16
    q = np.array([[-1.0, -0.5, 0.0, 0.5, 1.0, 32767 / 32768]], dtype=np.float32)
    assert helpers.pcm16_from_float(q).tolist() == [[-32768, -16384, 0, 16384, 32767, 32767]]
    import wave                                                  # the stdlib reader agrees on the container
    with wave.open(p) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 24000, 5000)


def test_read_wav_formats(tmp_path):
    import struct
    n, rate = 100, 8000
    x = np.linspace(-0.9, 0.9, n).astype(np.float32)

    def write(path, tag, bits, payload):
        hdr = b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE" + b"fmt " + struct.pack(
            "<IHHIIHH", 16, tag, 1, rate, rate * bits // 8, bits // 8, bits) + b"data" + struct.pack("<I", len(payload))
        open(path, "wb").write(hdr + payload)

    write(tmp_path / "f32.wav", 3, 32, x.astype("<f4").tobytes())
    y, r = helpers.read_wav(str(tmp_path / "f32.wav"))
    assert r == rate and np.array_equal(y[0], x)
    v24 = np.rint(x.astype(np.float64) * 8388607).astype(np.int32)
    b = bytearray()
    for v in v24:
        b += int(v & 0xFFFFFF).to_bytes(3, "little")
    write(tmp_path / "i24.wav", 1, 24, bytes(b))
    y, _ = helpers.read_wav(str(tmp_path / "i24.wav"))
    assert np.abs(y[0] - v24 / 8388608.0).max() < 1e-7
    write(tmp_path / "u8.wav", 1, 8, (np.rint(x * 127) + 128).astype(np.uint8).tobytes())
    y, _ = helpers.read_wav(str(tmp_path / "u8.wav"))
    assert np.abs(y[0] - np.rint(x * 127) / 128.0).max() < 1e-7
    with pytest.raises(ValueError):
        helpers.read_wav(__file__)


@pytest.mark.parametrize("orig,new", [(24000, 16000), (44100, 16000), (8000, 16000), (48000, 16000)])
def test_resample_matches_torchaudio(orig, new):
    ta = pytest.importorskip("torchaudio")
    g = torch.Generator().manual_seed(orig)
    x = torch.randn(2, 3 * orig // 10 + 17, generator=g)
    ref = ta.functional.resample(x, orig, new)
    out = helpers.resample(x, orig, new)
    assert out.shape == ref.shape == (2, math.ceil(new * x.shape[-1] / orig))
    assert (out - ref).abs().max().item() < 5e-5      # torchaudio builds the same kernel in float32, here in float64


def test_resample_tone_and_load_audio(tmp_path):
    rate, f0 = 48000, 1000.0
    t = np.arange(rate) / rate
    x = 0.5 * np.sin(2 * np.pi * f0 * t)
    p = str(tmp_path / "tone.wav")
    helpers.write_wav_pcm16(p, np.stack([x, x]), rate)           # stereo -> mono mix
    w = helpers.load_audio(p, 16000)
    assert w.shape == (1, 1, 16000) and w.dtype == torch.float32
    ref = 0.5 * np.sin(2 * np.pi * f0 * np.arange(16000) / 16000)
    assert np.abs(w[0, 0, 200:-200].numpy() - ref[200:-200]).max() < 2e-3
    with pytest.raises(RuntimeError):
        helpers.load_audio(str(tmp_path / "x.mp3"), 16000)
    assert helpers.find_audio_files(str(tmp_path)) == [p]


def test_bitstream_roundtrip_and_rate():
    g = np.random.default_rng(1)
    for frames in (0, 1, 7, 125, 375, 1001):
        c = g.integers(0, 2016, size=(8, frames), dtype=np.int32)
        blob = bitstream.pack_codes(torch.from_numpy(c))
        assert len(blob) == bitstream.HEADER.size + (frames * 88 + 7) // 8       # 11 bytes per 80 ms frame
        assert np.array_equal(bitstream.unpack_codes(blob), c)
    assert bitstream.bitrate_bps() == 1100.0
    edge = np.array([[0, 2015]] * 8, dtype=np.int32)
    assert np.array_equal(bitstream.unpack_codes(bitstream.pack_codes(edge)), edge)
    with pytest.raises(ValueError):
        bitstream.pack_codes(np.full((8, 2), 2016))
    with pytest.raises(ValueError):
        bitstream.unpack_codes(b"nope" + bytes(20))
    with pytest.raises(ValueError):
        bitstream.unpack_codes(bitstream.pack_codes(edge)[:-3])


def test_bitstream_known_answer():
    # one frame, codes 1, 2, 3, ...: LSB-first 11-bit fields
    c = np.arange(1, 9, dtype=np.int32)[:, None]
    payload = bitstream.pack_codes(c)[bitstream.HEADER.size:]
    word = sum(int(v) << (11 * i) for i, v in enumerate(c[:, 0]))
    assert payload == word.to_bytes(11, "little")

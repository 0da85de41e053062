"""End-to-end CLI run on the GPU (the counterpart of the reference's inference.py): wav files in, wav files + code
streams out."""
import os

import numpy as np
import pytest
import torch

from conftest import synthetic_wave
from simwhisper_codec_b200 import bitstream, inference
from simwhisper_codec_b200.utils import helpers

pytestmark = pytest.mark.gpu


def test_cli_directory_roundtrip(tmp_path, gen_params, sd_ex):
    ind, outd, cd = tmp_path / "in", tmp_path / "out", tmp_path / "codes"
    os.makedirs(ind / "sub")
    lens = {"a": 48123, "sub/b": 36000 * 3 // 2, "c": 16000 * 31}          # c spans two 30 s windows; b is 24 kHz stereo
    helpers.write_wav_pcm16(str(ind / "a.wav"), synthetic_wave(1, lens["a"]).numpy(), 16000)
    st = torch.stack([synthetic_wave(2, lens["sub/b"]), synthetic_wave(3, lens["sub/b"])]).numpy()
    helpers.write_wav_pcm16(str(ind / "sub" / "b.wav"), st, 24000)
    helpers.write_wav_pcm16(str(ind / "c.wav"), synthetic_wave(4, lens["c"]).numpy(), 16000)
    rc = inference.main(["--input_dir", str(ind), "--output_dir", str(outd), "--codes_dir", str(cd), "--random_init",
                         "--precision", "fp32", "--batch_size", "2"])
    assert rc == 0
    expect = {"a": 48123, "b": 36000, "c": 496000}                           # 16 kHz sample counts after resampling
    from simwhisper_codec_b200 import AudioCodec
    model = AudioCodec(gen_params, precision="fp32")
    model.load_state_dict(sd_ex)
    for name, n in expect.items():
        y, rate = helpers.read_wav(str(outd / f"{name}.wav"))
        codes = bitstream.unpack_codes(open(cd / f"{name}.swc", "rb").read())
        assert rate == 16000 and codes.shape == (8, n // 1280) and y.shape == (1, (n // 1280) * 1280)
        assert np.isfinite(y).all() and float(np.abs(y).max()) > 0
    # the stream holds exactly what encode() returns for the same (16-bit quantised) input
    w, _ = helpers.read_wav(str(ind / "a.wav"))
    ref = model.encode([torch.from_numpy(w[0])])["codes_list"][0].cpu().numpy()
    assert np.array_equal(bitstream.unpack_codes(open(cd / "a.swc", "rb").read()), ref)

"""Host-side packing (libswc swc_model_pack, no GPU) + closed-form restatements, checked against the
oracle: the kernels' arithmetic is emulated in float64 on the PACKED tables (tests/emulate.py)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, snr_db, synthetic_wave
from emulate import Emu
from oracle import port
from simwhisper_codec_b200 import AudioCodec, _lib


@pytest.fixture(scope="module")
def emu(gen_params, sd_ex):
    m = AudioCodec(gen_params, precision="fp32")
    m.load_state_dict(sd_ex, strict=True)
    return Emu(m.pack_preview())


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "swc.h")).read()
    declared = set(re.findall(r"\b(swc_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.swc_version() >= 100


def test_header_is_plain_c_and_links(tmp_path):
    """include/swc.h is the boundary a host in any language binds: it must compile as C99 (no C++ / torch types) and a
    plain-C program must link against libswc.so and drive the model lifecycle and the error path (examples/c_abi_check.c)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "c_abi_check")
    libdir = os.path.join(ROOT, "simwhisper_codec_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_check.c"), "-o", exe, os.path.join(libdir, "libswc.so"),
                    "-Wl,-rpath," + libdir], check=True, capture_output=True, timeout=120)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "state_dict" in r.stdout and r.stdout.strip().endswith("ok")


def test_compute_fails_loudly_without_gpu(gen_params):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = AudioCodec(gen_params)
    with pytest.raises(RuntimeError):
        m.inference_tokenize(torch.zeros(1, 1, 16000), torch.tensor([16000]))
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.swc_model_create(C.byref(h), 0) == 0
    assert lib.swc_model_finalize(h, 0) != 0          # no device -> error, no CPU path
    assert b"CUDA" in lib.swc_last_error() or b"state_dict" in lib.swc_last_error()
    lib.swc_model_destroy(h)


def test_fsq_constants(emu):
    c = emu.nat.packed("fsq.const").numpy()
    np.testing.assert_allclose(c[:4], [3.4965, 2.997, 2.4975, 2.4975], rtol=1e-6)
    np.testing.assert_allclose(c[8:12], [0.143982917, 0, 0.202918470, 0.202918470], atol=2e-8)
    k = load_golden("fsq_kat.npz")
    z = torch.from_numpy(k["z"])                     # (1,4,256): one group -> replicate to 8 groups
    lat = z.transpose(1, 2).repeat(1, 1, 8)
    dq, codes = emu.fsq(lat, torch.tensor([256]))
    assert torch.equal(codes[0], torch.from_numpy(k["idx"])[0])
    assert torch.equal(dq[0, :, :4].T, torch.from_numpy(k["dq"])[0])


def test_mel_tables_and_frontend(emu):
    t = load_golden("tables.npz")
    fb = emu.nat.packed("mel.fb.w").view(80, 208)[:, :201].numpy()
    np.testing.assert_allclose(fb, t["mel_filters"].T.astype(np.float32), rtol=0, atol=1e-9)
    g = load_golden("api_10s_ex.npz")
    w = synthetic_wave(1000, 160000)
    mel = emu.mel(w[None, :].double(), torch.tensor([160000]))
    assert np.abs(mel[0, :, :1008].numpy() - g["mel"]).max() < 1e-4        # north_star bound on mel


def test_forward_small_emulated(emu):
    g = load_golden("forward_small_ex.npz")
    mel, lens = torch.from_numpy(g["mel"]).double(), torch.from_numpy(g["mel_lens"])
    enc_cl, enc_lens = emu.encoder(mel, lens)
    T = g["enc"].shape[2]
    assert np.abs(enc_cl[:, :T].transpose(1, 2).numpy() - g["enc"]).max() < 2e-5
    lat, code_lens = emu.downsample(enc_cl, enc_lens)
    assert np.abs(lat.transpose(1, 2).numpy() - g["latent"]).max() < 5e-5
    dq, codes = emu.fsq(lat, code_lens)
    flips = (codes != torch.from_numpy(g["codes"])).sum().item()
    assert flips == 0
    h = emu.upsample(dq)
    assert np.abs(h.transpose(1, 2).numpy() - g["up"]).max() < 5e-5
    mel_dec = emu.decoder(h, code_lens * 4)
    assert np.abs(mel_dec[..., :80].transpose(1, 2).numpy() - g["dec"]).max() < 5e-5
    assert float(mel_dec[..., 80:].abs().max()) == 0.0
    wav = emu.vocos(mel_dec)
    assert snr_db(torch.from_numpy(g["audio"])[:, 0], wav) > 80.0


def test_bf16x3_split_arithmetic():
    """The operand split of the bf16x3 mode (pack.cu / split_bf16_planes): x = hi + lo up to 2^-17 |x|, and the three
    products hi*hi + hi*lo + lo*hi with fp32 accumulation reproduce an fp32 contraction to ~1e-5 of sum |a||w|."""
    import numpy as np
    import torch
    g = torch.Generator().manual_seed(0)
    a = torch.randn(64, 768, generator=g)
    w = torch.randn(96, 768, generator=g) * 0.05
    def split(x):
        hi = x.bfloat16().float()
        lo = (x - hi).bfloat16().float()
        return hi, lo
    ah, al = split(a)
    wh, wl = split(w)
    assert float(((a - ah - al).abs() / a.abs().clamp_min(1e-30)).max()) <= 2.0 ** -16
    ref = a.double() @ w.double().T
    got = (ah @ wh.T + ah @ wl.T + al @ wh.T).double()
    scale = (a.abs().double() @ w.abs().double().T)
    assert float(((got - ref).abs() / scale).max()) < 2e-5
    plain = (ah @ wh.T).double()                       # single bf16 product for contrast: ~2^-9
    assert float(((plain - ref).abs() / scale).max()) > 1e-4

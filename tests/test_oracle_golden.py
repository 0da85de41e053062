"""Pin the CPU oracle (oracle/port.py) and the deterministic weights against fixtures produced by
the real reference (tests/golden/make_goldens.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, snr_db, synthetic_wave
from oracle import port
from simwhisper_codec_b200 import weights as W


def test_schema_matches_reference(gen_params):
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_schema.json")))
    ours = W.state_dict_schema(gen_params)
    assert list(ours.keys()) == list(ref.keys())
    for k, (shape, dtype, _) in ours.items():
        assert [list(shape), str(dtype)] == ref[k], k
    assert len(ours) == 711


def test_weight_digests(sd_ex, sd_plain):
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
    assert W.state_dict_digest(sd_ex) == meta["digest_ex"]
    assert W.state_dict_digest(sd_plain) == meta["digest_plain"]


def test_constant_tables():
    t = load_golden("tables.npz")
    np.testing.assert_allclose(port.mel_filterbank(), t["mel_filters"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(W.kaiser_sinc_taps().astype(np.float32), t["aa_filter"], rtol=0, atol=6e-8)
    np.testing.assert_allclose(W.hann_periodic(640).astype(np.float32), t["istft_window"], rtol=0, atol=5e-7)
    np.testing.assert_allclose(W.hann_periodic(400).astype(np.float32), t["hann400"], rtol=0, atol=5e-7)


def test_fsq_known_answers():
    k = load_golden("fsq_kat.npz")
    z = torch.from_numpy(k["z"])
    dq, idx = port.fsq_encode(z, torch.tensor([z.shape[-1]]))
    assert torch.equal(idx, torch.from_numpy(k["idx"]))
    assert torch.equal(dq, torch.from_numpy(k["dq"]))
    assert idx[0, 0, :6].tolist() == [1204, 2006, 401, 2015, 0, 581]
    dec = port.fsq_decode(idx, torch.tensor([z.shape[-1]]))
    assert torch.equal(dec, torch.from_numpy(k["dec"]))
    assert torch.equal(dec, dq)
    # masking beyond the length
    dq2, idx2 = port.fsq_encode(z, torch.tensor([100]))
    assert int(idx2[..., 100:].abs().max()) == 0 and float(dq2[..., 100:].abs().max()) == 0.0


@pytest.mark.parametrize("tag", ["ex", "plain"])
def test_forward_small(tag, sd_ex, sd_plain):
    sd = sd_ex if tag == "ex" else sd_plain
    g = load_golden(f"forward_small_{tag}.npz")
    mel, lens = torch.from_numpy(g["mel"]), torch.from_numpy(g["mel_lens"])
    with torch.inference_mode():
        enc, el = port.encoder(sd, mel, lens)
        lat, ll = port.downsample(sd, enc, el)
        zq, codes = port.fsq_encode(lat, ll)
        up, ul = port.upsample(sd, zq, ll)
        dec, dl = port.decoder(sd, up, ul)
        y, yl = port.vocos(sd, dec, dl)
    # same ATen ops as the reference on the same machine: expect (near) bit equality
    assert torch.allclose(enc, torch.from_numpy(g["enc"]), atol=2e-6, rtol=0)
    assert torch.allclose(lat, torch.from_numpy(g["latent"]), atol=1e-5, rtol=0)  # weight-norm formula: few ulp
    assert torch.equal(codes, torch.from_numpy(g["codes"]))
    assert torch.equal(zq, torch.from_numpy(g["zq"]))
    assert torch.allclose(up, torch.from_numpy(g["up"]), atol=1e-5, rtol=0)
    assert torch.allclose(dec, torch.from_numpy(g["dec"]), atol=1e-5, rtol=0)
    assert snr_db(torch.from_numpy(g["audio"]), y) > 100.0
    assert yl.tolist() == g["audio_lengths"].tolist()


def test_api_10s(sd_ex):
    g = load_golden("api_10s_ex.npz")
    w = synthetic_wave(1000, 160000)
    trace = {}
    with torch.inference_mode():
        r = port.tokenize(sd_ex, w[None, None, :], torch.tensor([160000]), trace=trace)
        codes = port.encode(sd_ex, [w])
        wav = port.decode(sd_ex, codes)
    assert np.abs(trace["mel"][0, :, :1008].numpy() - g["mel"]).max() < 1e-5
    assert np.abs(trace["mel"][0, :, 1008:].numpy()[:, ::97] - g["mel_tail"]).max() < 1e-5
    assert int(trace["mel_lens"][0]) == 1000 and int(r["codes_lengths"][0]) == 125
    assert np.abs(trace["latent"][0, :, :125].numpy() - g["latent"]).max() < 2e-5
    assert tuple(codes[0].shape) == (8, 125)
    assert torch.equal(codes[0], torch.from_numpy(g["codes"]))
    assert wav[0].shape[0] == 160000
    assert snr_db(torch.from_numpy(g["wav"]), wav[0]) > 100.0


def test_api_batch_windows(sd_ex):
    """Variable-length batch with a 50 s item: 3 windows, keep-first stitching, decode T' = batch max."""
    g = load_golden("api_batch_ex.npz")
    lens = g["lens"].tolist()
    wavs = [synthetic_wave(2000 + i, n) for i, n in enumerate(lens)]
    with torch.inference_mode():
        codes = port.encode(sd_ex, wavs)
        assert [tuple(c.shape) for c in codes] == [(8, 37), (8, 625), (8, 285)]
        for i, c in enumerate(codes):
            assert torch.equal(c, torch.from_numpy(g[f"codes{i}"])), i
        wav = port.decode(sd_ex, codes)
    assert [len(x) for x in wav] == g["wav_len"].tolist() == [47360, 800000, 364800]
    assert snr_db(torch.from_numpy(g["wav0"]), wav[0]) > 100
    assert snr_db(torch.from_numpy(g["wav1_head"]), wav[1][:64000]) > 100
    assert snr_db(torch.from_numpy(g["wav1_seam"]), wav[1][310000:330000]) > 100
    assert snr_db(torch.from_numpy(g["wav1_tail"]), wav[1][-32000:]) > 100
    assert snr_db(torch.from_numpy(g["wav2_seam"]), wav[2][312000:328000]) > 100


def test_encoder_hidden_states(sd_ex):
    """output_hidden_states=True (reference modules.py:344-371) against the reference's own 13 tensors."""
    g = load_golden("hidden_small_ex.npz")
    f = load_golden("forward_small_ex.npz")
    mel, lens = torch.from_numpy(f["mel"]), torch.from_numpy(f["mel_lens"])
    with torch.inference_mode():
        out, ol, hs = port.encoder(sd_ex, mel, lens, output_hidden_states=True)
    assert len(hs) == 13 and ol.tolist() == g["out_len"].tolist()
    assert torch.allclose(out[:, ::8], torch.from_numpy(g["out"]), atol=2e-6, rtol=0)
    for i, h in enumerate(hs):
        assert torch.allclose(h[:, ::8], torch.from_numpy(g["hidden"][i]), atol=5e-6, rtol=0), i
    assert float(hs[3][1, :, 68:].abs().max()) == 0.0           # item 1 has 137 // 2 = 68 valid tokens: the rest is masked


def test_outlier_weights_forward(gen_params):
    """Whisper-like outlier-channel weights: the port follows the reference there too (fp32 vs fp32)."""
    g = load_golden("outlier_ex.npz")
    sd = W.random_state_dict(gen_params, seed=0, exercise=True, outlier_gain=50.0)
    assert W.state_dict_digest(sd) == str(g["digest"])
    mel, lens = torch.from_numpy(g["mel"]), torch.from_numpy(g["mel_lens"])
    with torch.inference_mode():
        enc, el = port.encoder(sd, mel, lens)
        lat, ll = port.downsample(sd, enc, el)
        zq, codes = port.fsq_encode(lat, ll)
    assert torch.allclose(enc, torch.from_numpy(g["enc"]), atol=2e-5, rtol=0)
    flips = (codes != torch.from_numpy(g["codes"])).float().mean()
    assert float(flips) < 1e-3, float(flips)
    rms = g["channel_rms"]
    assert rms[list(W.OUTLIER_CHANNELS)].min() > 20 * np.median(rms)      # the stress is what it says: >20x outlier channels


def test_vocos_receptive_field_bounds_the_packed_form(sd_ex):
    """decode() with host-known lengths computes only need = min(Tv, 8 len + 80) Vocos frames per window and treats the cut
    as a sequence end (pipeline.cu::detokenize_chain).  In the reference arithmetic (float64 here) the first 1280 len samples of
    a window must then not depend on what lies beyond the cut: 3 frames per depthwise convolution and for the embedding
    (25 x 3 = 75) + 1 for the overlap-add = 76 < 80.  A halo of 4 frames must NOT be enough (the cut does reach valid samples
    when it sits too close; with these weights the influence decays by ~10^3 per 8 frames)."""
    sd = port.cast_sd(sd_ex, torch.float64)
    g = torch.Generator().manual_seed(11)
    Tv = 232
    x = torch.randn(1, 80, Tv, generator=g, dtype=torch.float64)
    with torch.inference_mode():
        full = port.vocos(sd, x, torch.tensor([Tv]))[0][0, 0]
        for code_len in (1, 9, 18):
            need = min(Tv, 8 * code_len + 80)
            cut = port.vocos(sd, x[:, :, :need].contiguous(), torch.tensor([need]))[0][0, 0]
            n = 1280 * code_len
            assert (cut[:n] - full[:n]).abs().max().item() <= 1e-10 * full.abs().max().item(), code_len
        short = port.vocos(sd, x[:, :, : 8 * 9 + 4].contiguous(), torch.tensor([8 * 9 + 4]))[0][0, 0]
        assert (short[: 1280 * 9] - full[: 1280 * 9]).abs().max().item() > 1e-3 * full.abs().max().item()


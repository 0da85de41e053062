"""Generate the committed golden fixtures by running the REAL reference (read-only tree at
/root/reference, or $SIMWHISPER_REF) on CPU with the deterministic weights of
`simwhisper_codec_b200.weights`.  Run from the repo root in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_goldens.py

The reference tree does not exist on the GPU box; tests only read the .npz/.json files written here.
Inputs are regenerated in the tests from the same seeds (`synthetic_wave` below is duplicated in
tests/conftest.py on purpose so the tests do not import this script's reference dependency).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SIMWHISPER_REF", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from simwhisper_codec_b200.weights import random_state_dict, state_dict_digest  # noqa: E402

import warnings  # noqa: E402

warnings.filterwarnings("ignore")
from audiocodec.model import AudioCodec  # noqa: E402  (the reference)
from audiocodec.nn.quantizer import FiniteScalarQuantizer  # noqa: E402


def synthetic_wave(seed: int, n: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(n, generator=g)).clamp_(-1, 1)


def main():
    torch.set_num_threads(os.cpu_count())
    cfg = yaml.safe_load(open(os.path.join(ROOT, "simwhisper_codec_b200/config/SimWhisperCodec.yaml")))
    gp = cfg["generator_params"]
    ref_cfg = yaml.safe_load(open(os.path.join(REF, "config/SimWhisperCodec.yaml")))
    assert ref_cfg == cfg, "our YAML must carry the same keys/values as the reference's"

    t0 = time.time()
    model = AudioCodec(gp).eval()
    schema = {k: [list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()}
    json.dump(schema, open(os.path.join(HERE, "state_dict_schema.json"), "w"), indent=0)
    print("reference built", time.time() - t0)

    # ---- small constant tables the kernels re-derive
    fe = model.feature_extractor
    np.savez_compressed(
        os.path.join(HERE, "tables.npz"),
        mel_filters=np.asarray(fe.mel_filters, dtype=np.float64),
        aa_filter=model.downsample.res_blocks[0].block[0].upsample.filter.numpy().reshape(-1),
        istft_window=model.vocos.head.istft.window.numpy(),
        hann400=torch.hann_window(400).numpy(),
    )

    # ---- FSQ known answers straight from the reference quantizer
    fsq = FiniteScalarQuantizer(num_levels=[8, 7, 6, 6], eps=1e-3)
    g = torch.Generator().manual_seed(7)
    z = torch.cat([
        torch.tensor([[0, 0, 0, 0], [1, 1, 1, 1], [-1, -1, -1, -1], [10, 10, 10, 10], [-10, -10, -10, -10],
                      [.3, -.2, .7, -.9]], dtype=torch.float32),
        torch.randn(250, 4, generator=g) * 1.5,
    ]).T[None]                                   # (1,4,256)
    dq, idx = fsq(z, torch.tensor([256]))
    np.savez_compressed(os.path.join(HERE, "fsq_kat.npz"), z=z.numpy(), dq=dq.numpy(), idx=idx.numpy(),
                        dec=fsq.decode(idx, torch.tensor([256])).numpy())

    meta = {}
    for tag, exercise in (("ex", True), ("plain", False)):
        sd = random_state_dict(gp, seed=0, exercise=exercise)
        meta[f"digest_{tag}"] = state_dict_digest(sd)
        model.load_state_dict(sd, strict=True)

        # ---- single forward at small T through AudioCodec.forward, with module-level captures
        cap = {}
        hooks = []
        for name in ("acoustic_encoder", "downsample", "quantizer", "upsample", "acoustic_decoder", "vocos"):
            hooks.append(getattr(model, name).register_forward_hook(
                lambda m, i, o, name=name: cap.__setitem__(name, o)))
        gm = torch.Generator().manual_seed(11)
        mel = torch.randn(2, 80, 200, generator=gm) * 0.5
        mel_lens = torch.tensor([200, 137])
        with torch.inference_mode():
            out = model({"mel_features": mel, "mel_lens": mel_lens})
        for h in hooks:
            h.remove()
        np.savez_compressed(
            os.path.join(HERE, f"forward_small_{tag}.npz"),
            mel=mel.numpy(), mel_lens=mel_lens.numpy(),
            enc=cap["acoustic_encoder"][0].numpy(), latent=cap["downsample"][0].numpy(),
            zq=cap["quantizer"][0].numpy(), codes=cap["quantizer"][1].numpy(),
            up=cap["upsample"][0].numpy(), dec=cap["acoustic_decoder"][0].numpy(),
            audio=out["reconstructed_audio"].numpy(), audio_lengths=out["audio_lengths"].numpy())
        print(tag, "forward_small done; distinct codes", len(np.unique(cap["quantizer"][1].numpy())), time.time() - t0)
        if not exercise:
            continue

        # ---- BASELINE.json configs[0]: one 10 s utterance through encode()/decode() on CPU
        w = synthetic_wave(1000, 160000)
        cap = {}
        hooks = [model.acoustic_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("enc", (i, o))),
                 model.downsample.register_forward_hook(lambda m, i, o: cap.__setitem__("lat", o)),
                 model.acoustic_decoder.register_forward_hook(lambda m, i, o: cap.__setitem__("dec", o))]
        with torch.inference_mode():
            codes = model.encode([w], overlap_seconds=10, device=torch.device("cpu"))["codes_list"]
            wav = model.decode(codes, overlap_seconds=10, device=torch.device("cpu"))["syn_wav_list"]
        for h in hooks:
            h.remove()
        mel10 = cap["enc"][0][0]
        np.savez_compressed(
            os.path.join(HERE, "api_10s_ex.npz"),
            codes=codes[0].numpy(), wav=wav[0].numpy(),
            mel=mel10[0, :, :1008].numpy(), mel_tail=mel10[0, :, 1008:].numpy()[:, ::97],
            enc=cap["enc"][1][0][0, ::16, :500].numpy(),
            latent=cap["lat"][0][0, :, :125].numpy(),
            dec=cap["dec"][0][0, ::8].numpy())
        print("api_10s done", codes[0].shape, wav[0].shape, "distinct", len(torch.unique(codes[0])), time.time() - t0)

        # ---- variable-length batch incl. a 50 s item (3 windows) through the API
        lens = [48123, 800000, 365000]
        wavs = [synthetic_wave(2000 + i, n) for i, n in enumerate(lens)]
        with torch.inference_mode():
            codes = model.encode(wavs, overlap_seconds=10, device=torch.device("cpu"))["codes_list"]
            wav = model.decode(codes, overlap_seconds=10, device=torch.device("cpu"))["syn_wav_list"]
        np.savez_compressed(
            os.path.join(HERE, "api_batch_ex.npz"), lens=np.asarray(lens),
            **{f"codes{i}": c.numpy() for i, c in enumerate(codes)},
            # waveforms are large: keep item 0 whole, and strided/segment views of the others
            wav0=wav[0].numpy(), wav1_head=wav[1][:64000].numpy(), wav1_seam=wav[1][310000:330000].numpy(),
            wav1_tail=wav[1][-32000:].numpy(), wav2_head=wav[2][:32000].numpy(), wav2_seam=wav[2][312000:328000].numpy(),
            wav_len=np.asarray([len(x) for x in wav]),
            wav_rms=np.asarray([float(x.double().pow(2).mean().sqrt()) for x in wav]))
        print("api_batch done", [tuple(c.shape) for c in codes], time.time() - t0)

    json.dump(meta, open(os.path.join(HERE, "meta.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

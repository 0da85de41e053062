"""State-dict schema of the SimWhisper-Codec generator and a deterministic random initialiser.

The trained checkpoint is not available offline, so parity and throughput are measured on
random-init weights of the named architecture.  The reference builds its modules under torch's
global RNG (and builds the encoder twice, reference audiocodec/model.py:40-42), which does not
travel to a box without the reference tree.  This module therefore defines the *schema* (the 711
keys/shapes/dtypes of `AudioCodec.state_dict()`, reference audiocodec/model.py:40-57 and the
module constructors it calls) and fills it from a counter-free, per-key seeded numpy PCG64 stream,
so that this container (where the reference runs and goldens are made) and the GPU box build
bit-identical tensors.

Two flavours:
  * exercise=False : distributions equal to what the reference constructors produce
                     (LayerNorm = 1/0, Snake alpha/beta = 0, Vocos biases = 0, gamma = 1/24, ...).
  * exercise=True  : every learnable tensor is perturbed away from its trivial value so that a
                     kernel that ignores e.g. a LayerNorm bias or a Snake beta fails parity, and
                     `downsample.to_latent.weight_g` is scaled by `latent_gain` so the FSQ code
                     space is actually exercised.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------
# fixed buffers
# ----------------------------------------------------------------------------------------------


def sinusoid_table(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    """Whisper positional table (unused by the acoustic encoder/decoder but part of the state
    dict; reference audiocodec/nn/modules.py:52-58)."""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2, dtype=np.float32)).astype(np.float32)
    t = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def kaiser_sinc_taps(cutoff: float = 0.25, half_width: float = 0.3, taps: int = 12) -> np.ndarray:
    """12-tap Kaiser-windowed sinc low-pass used by the anti-aliased activations
    (reference audiocodec/nn/alias_free_torch/filter.py:25-54).  float64 result."""
    half = taps // 2
    delta_f = 4.0 * half_width
    att = 2.285 * (half - 1) * math.pi * delta_f + 7.95
    if att > 50.0:
        beta = 0.1102 * (att - 8.7)
    elif att >= 21.0:
        beta = 0.5842 * (att - 21.0) ** 0.4 + 0.07886 * (att - 21.0)
    else:
        beta = 0.0
    n = np.arange(taps, dtype=np.float64)
    # symmetric (non-periodic) Kaiser window
    window = np.i0(beta * np.sqrt(1.0 - ((n - (taps - 1) / 2.0) / ((taps - 1) / 2.0)) ** 2)) / np.i0(beta)
    time = np.arange(-half, half, dtype=np.float64) + 0.5 if taps % 2 == 0 else n - half
    filt = 2.0 * cutoff * window * np.sinc(2.0 * cutoff * time)
    return filt / filt.sum()


def hann_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n) (periodic), float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


# ----------------------------------------------------------------------------------------------
# schema
# ----------------------------------------------------------------------------------------------
# init kinds:  ("uniform", bound) ("normal", std, clip) ("const", v) ("ln_w",) ("ln_b",)
#              ("wn_g", key_of_v) ("snake",) ("gamma", v) ("buffer", ndarray)


def _transformer_layer(prefix: str, d: int, ffn: int, out: "OrderedDict"):
    b = 1.0 / math.sqrt(d)
    bf = 1.0 / math.sqrt(ffn)
    out[prefix + "self_attn.k_proj.weight"] = ((d, d), torch.float32, ("uniform", b))
    out[prefix + "self_attn.v_proj.weight"] = ((d, d), torch.float32, ("uniform", b))
    out[prefix + "self_attn.v_proj.bias"] = ((d,), torch.float32, ("uniform", b))
    out[prefix + "self_attn.q_proj.weight"] = ((d, d), torch.float32, ("uniform", b))
    out[prefix + "self_attn.q_proj.bias"] = ((d,), torch.float32, ("uniform", b))
    out[prefix + "self_attn.out_proj.weight"] = ((d, d), torch.float32, ("uniform", b))
    out[prefix + "self_attn.out_proj.bias"] = ((d,), torch.float32, ("uniform", b))
    out[prefix + "self_attn_layer_norm.weight"] = ((d,), torch.float32, ("ln_w",))
    out[prefix + "self_attn_layer_norm.bias"] = ((d,), torch.float32, ("ln_b",))
    out[prefix + "fc1.weight"] = ((ffn, d), torch.float32, ("uniform", b))
    out[prefix + "fc1.bias"] = ((ffn,), torch.float32, ("uniform", b))
    out[prefix + "fc2.weight"] = ((d, ffn), torch.float32, ("uniform", bf))
    out[prefix + "fc2.bias"] = ((d,), torch.float32, ("uniform", bf))
    out[prefix + "final_layer_norm.weight"] = ((d,), torch.float32, ("ln_w",))
    out[prefix + "final_layer_norm.bias"] = ((d,), torch.float32, ("ln_b",))


def _wn_conv(prefix: str, cout: int, cin: int, k: int, out: "OrderedDict"):
    b = 1.0 / math.sqrt(cin * k)
    out[prefix + "bias"] = ((cout,), torch.float32, ("small_bias", b))
    out[prefix + "weight_g"] = ((cout, 1, 1), torch.float32, ("wn_g", prefix + "weight_v"))
    out[prefix + "weight_v"] = ((cout, cin, k), torch.float32, ("uniform", b))


def _res_blocks(prefix: str, hidden: int, out: "OrderedDict"):
    taps = kaiser_sinc_taps().astype(np.float32).reshape(1, 1, 12)
    for i in range(3):
        p = f"{prefix}res_blocks.{i}.block."
        for j in (0, 2):
            out[p + f"{j}.act.alpha"] = ((hidden,), torch.float32, ("snake",))
            out[p + f"{j}.act.beta"] = ((hidden,), torch.float32, ("snake",))
            out[p + f"{j}.upsample.filter"] = ((1, 1, 12), torch.float32, ("buffer", taps))
            out[p + f"{j}.downsample.lowpass.filter"] = ((1, 1, 12), torch.float32, ("buffer", taps))
            if j == 0:
                _wn_conv(p + "1.", hidden, hidden, 7, out)
        _wn_conv(p + "3.", hidden, hidden, 1, out)


def state_dict_schema(gp: dict) -> "OrderedDict[str, Tuple[tuple, torch.dtype, tuple]]":
    """Ordered key -> (shape, dtype, init spec) for `generator_params` `gp`."""
    out: "OrderedDict[str, Tuple[tuple, torch.dtype, tuple]]" = OrderedDict()
    enc, dec, dn, up, q, vo = (gp[k] for k in ("acoustic_encoder", "acoustic_decoder", "downsample",
                                               "upsample", "quantizer", "vocos"))
    # --- encoder (reference modules.py:237-285)
    d, mel, ks = enc["d_model"], enc["num_mel_bins"], enc["kernel_size"]
    npos = (enc["max_audio_seconds"] * enc["sampling_rate"] // enc["hop_length"]) // enc["stride_size"]
    out["acoustic_encoder.positional_embedding"] = ((npos, d), torch.float32, ("buffer", sinusoid_table(npos, d)))
    b1, b2 = 1.0 / math.sqrt(mel * ks), 1.0 / math.sqrt(d * ks)
    out["acoustic_encoder.conv1.weight"] = ((d, mel, ks), torch.float32, ("uniform", b1))
    out["acoustic_encoder.conv1.bias"] = ((d,), torch.float32, ("uniform", b1))
    out["acoustic_encoder.conv2.weight"] = ((d, d, ks), torch.float32, ("uniform", b2))
    out["acoustic_encoder.conv2.bias"] = ((d,), torch.float32, ("uniform", b2))
    for i in range(enc["encoder_layers"]):
        _transformer_layer(f"acoustic_encoder.layers.{i}.", d, enc["encoder_ffn_dim"], out)
    out["acoustic_encoder.layer_norm.weight"] = ((d,), torch.float32, ("ln_w",))
    out["acoustic_encoder.layer_norm.bias"] = ((d,), torch.float32, ("ln_b",))
    # --- downsampler (modules.py:488-517)
    hid, s = dn["hidden_dim"], dn["stack_factor"]
    _wn_conv("downsample.in_proj.", hid, dn["in_dim"] * s, 1, out)
    _res_blocks("downsample.", hid, out)
    _wn_conv("downsample.to_latent.", dn["latent_dim"], hid, 1, out)
    # --- quantizer (quantizer.py:59-71)
    levels = list(q["num_levels_per_group"])
    base = np.cumprod([1] + levels[:-1]).astype(np.int32).reshape(1, -1, 1)
    lv = np.asarray(levels, dtype=np.int32).reshape(1, -1, 1)
    for g in range(q["num_groups"]):
        out[f"quantizer.fsqs.{g}.dim_base_index"] = ((1, len(levels), 1), torch.int32, ("buffer", base))
        out[f"quantizer.fsqs.{g}.num_levels"] = ((1, len(levels), 1), torch.int32, ("buffer", lv))
    # --- upsampler (modules.py:569-599)
    hid = up["hidden_dim"]
    _wn_conv("upsample.from_latent.", hid, up["latent_dim"], 1, out)
    _res_blocks("upsample.", hid, out)
    _wn_conv("upsample.to_stacked.", up["out_dim"] * up["stack_factor"], hid, 1, out)
    # --- decoder (modules.py:381-435)
    d, mel, ks = dec["d_model"], dec["num_mel_bins"], dec["kernel_size"]
    npos = (dec["max_audio_seconds"] * dec["sampling_rate"] // dec["hop_length"]) // dec["stride_size"]
    out["acoustic_decoder.positional_embedding"] = ((npos, d), torch.float32, ("buffer", sinusoid_table(npos, d)))
    # ConvTranspose1d weights are [in, out, k]; torch's fan_in for them is out*k
    bd1, bd2 = 1.0 / math.sqrt(d * ks), 1.0 / math.sqrt(mel * ks)
    out["acoustic_decoder.deconv1.weight"] = ((d, d, ks), torch.float32, ("uniform", bd1))
    out["acoustic_decoder.deconv1.bias"] = ((d,), torch.float32, ("uniform", bd1))
    out["acoustic_decoder.deconv2.weight"] = ((d, mel, ks), torch.float32, ("uniform", bd2))
    out["acoustic_decoder.deconv2.bias"] = ((mel,), torch.float32, ("uniform", bd2))
    for i in range(dec["decoder_layers"]):
        _transformer_layer(f"acoustic_decoder.layers.{i}.", d, dec["decoder_ffn_dim"], out)
    out["acoustic_decoder.layer_norm.weight"] = ((d,), torch.float32, ("ln_w",))
    out["acoustic_decoder.layer_norm.bias"] = ((d,), torch.float32, ("ln_b",))
    # --- Vocos (modules.py:1461-1490, 1052-1062, 819-829)
    dim, inter, cin, nl = vo["dim"], vo["intermediate_dim"], vo["input_channels"], vo["num_layers"]
    tn = ("normal", 0.02, 0.08)
    out["vocos.backbone.embed.weight"] = ((dim, cin, 7), torch.float32, tn)
    out["vocos.backbone.embed.bias"] = ((dim,), torch.float32, ("small_bias", 0.02))
    out["vocos.backbone.norm.weight"] = ((dim,), torch.float32, ("ln_w",))
    out["vocos.backbone.norm.bias"] = ((dim,), torch.float32, ("ln_b",))
    for i in range(nl):
        p = f"vocos.backbone.convnext.{i}."
        out[p + "gamma"] = ((dim,), torch.float32, ("gamma", 1.0 / nl))
        out[p + "dwconv.weight"] = ((dim, 1, 7), torch.float32, tn)
        out[p + "dwconv.bias"] = ((dim,), torch.float32, ("small_bias", 0.02))
        out[p + "norm.weight"] = ((dim,), torch.float32, ("ln_w",))
        out[p + "norm.bias"] = ((dim,), torch.float32, ("ln_b",))
        out[p + "pwconv1.weight"] = ((inter, dim), torch.float32, tn)
        out[p + "pwconv1.bias"] = ((inter,), torch.float32, ("small_bias", 0.02))
        out[p + "pwconv2.weight"] = ((dim, inter), torch.float32, tn)
        out[p + "pwconv2.bias"] = ((dim,), torch.float32, ("small_bias", 0.02))
    out["vocos.backbone.final_layer_norm.weight"] = ((dim,), torch.float32, ("ln_w",))
    out["vocos.backbone.final_layer_norm.bias"] = ((dim,), torch.float32, ("ln_b",))
    nout = vo["n_fft"] + 2
    bh = 1.0 / math.sqrt(dim)
    out["vocos.head.out.weight"] = ((nout, dim), torch.float32, ("uniform", bh))
    out["vocos.head.out.bias"] = ((nout,), torch.float32, ("uniform", bh))
    out["vocos.head.istft.window"] = ((vo["n_fft"],), torch.float32,
                                      ("buffer", hann_periodic(vo["n_fft"]).astype(np.float32)))
    return out


# ----------------------------------------------------------------------------------------------
# deterministic fill
# ----------------------------------------------------------------------------------------------


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))


def _uniform(rng: np.random.Generator, shape, bound: float) -> np.ndarray:
    # float64 draws mapped to (-bound, bound) then rounded once to float32
    return ((rng.random(shape) * 2.0 - 1.0) * bound).astype(np.float32)


OUTLIER_CHANNELS = (41, 302, 577, 760)


def random_state_dict(gp: dict, seed: int = 0, exercise: bool = True,
                      latent_gain: float = 4.0, outlier_gain: float = 0.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic random-init state dict with the reference's key schema (CPU tensors).

    outlier_gain > 0: Whisper-like outlier stress.  Trained Whisper encoders carry a few residual-stream channels whose
    magnitude is ~50x the rest; random init has none, so bf16 accuracy would only ever be tested on benign statistics.
    The rows of every tensor that WRITES the 768-wide residual stream (conv2, out_proj, fc2 of both transformer stacks)
    are scaled by `outlier_gain` for OUTLIER_CHANNELS, which puts those channels of h at ~gain x the others."""
    schema = state_dict_schema(gp)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    arrays: Dict[str, np.ndarray] = {}
    deferred = []
    for key, (shape, dtype, spec) in schema.items():
        kind = spec[0]
        rng = _rng(seed, key)
        if kind == "uniform":
            a = _uniform(rng, shape, spec[1])
        elif kind == "normal":
            a = np.clip(rng.standard_normal(shape) * spec[1], -spec[2], spec[2]).astype(np.float32)
        elif kind == "small_bias":
            a = _uniform(rng, shape, spec[1]) if exercise else np.zeros(shape, np.float32)
        elif kind == "ln_w":
            a = (1.0 + 0.1 * (rng.random(shape) * 2 - 1)).astype(np.float32) if exercise else np.ones(shape, np.float32)
        elif kind == "ln_b":
            a = _uniform(rng, shape, 0.05) if exercise else np.zeros(shape, np.float32)
        elif kind == "snake":
            a = _uniform(rng, shape, 0.3) if exercise else np.zeros(shape, np.float32)
        elif kind == "gamma":
            a = (spec[1] * (1.0 + 0.5 * (rng.random(shape) * 2 - 1))).astype(np.float32) if exercise \
                else np.full(shape, spec[1], np.float32)
        elif kind == "buffer":
            a = np.ascontiguousarray(spec[1]).reshape(shape)
        elif kind == "wn_g":
            deferred.append((key, shape, spec[1]))
            arrays[key] = None  # keep insertion order
            continue
        else:  # pragma: no cover
            raise ValueError(kind)
        arrays[key] = a
    for key, shape, vkey in deferred:
        v = arrays[vkey].astype(np.float64)
        g = np.sqrt((v * v).sum(axis=(1, 2), keepdims=True))
        if exercise:
            g = g * (1.0 + 0.1 * (_rng(seed, key).random(shape) * 2 - 1))
            if key == "downsample.to_latent.weight_g":
                g = g * latent_gain
        arrays[key] = g.astype(np.float32)
    if outlier_gain > 0:
        ch = list(OUTLIER_CHANNELS)
        for key in arrays:
            leaf = key.split(".")
            writes_h = (key == "acoustic_encoder.conv2.weight" or
                        (leaf[0] in ("acoustic_encoder", "acoustic_decoder") and leaf[1] == "layers" and
                         (leaf[3:5] == ["self_attn", "out_proj"] or leaf[3] == "fc2")))
            if writes_h:
                arrays[key] = arrays[key].copy()
                arrays[key][ch] = arrays[key][ch] * np.float32(outlier_gain)
    for key, (shape, dtype, _) in schema.items():
        t = torch.from_numpy(np.ascontiguousarray(arrays[key]))
        assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, (key, t.shape, t.dtype)
        sd[key] = t
    return sd


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    """Order-independent CRC of the raw bytes of every tensor (pins weights in goldens)."""
    acc = 0
    for k in sorted(sd):
        t = sd[k].detach().cpu().contiguous()
        acc = zlib.crc32(k.encode(), acc)
        acc = zlib.crc32(t.numpy().tobytes(), acc)
    return f"{acc:08x}"

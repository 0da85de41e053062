"""CPU oracle for the SimWhisper-Codec hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (state-dict in, tensors out) restatement of the reference's batched
`wav -> log-mel -> encoder -> downsample -> FSQ -> upsample -> decoder -> Vocos/iSTFT -> wav`
forward.  It is imported only by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`; the product package never imports it.

Parity status: PINNED.  The reference is pure Python and imports in the build container, so
`tests/golden/make_goldens.py` runs the real `audiocodec.model.AudioCodec` (from /root/reference)
on the deterministic weights of `simwhisper_codec_b200.weights.random_state_dict` and commits its
outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this port against them.
The reference itself has no tests or golden vectors.

Third-party arithmetic the reference delegates to (not under /root/reference, restated here):
  * transformers.audio_utils.mel_filter_bank (pinned transformers==4.53.3) -> `mel_filterbank`
  * transformers SequenceFeatureExtractor.pad (right zero-padding to 480000)  -> `log_mel`
  * torch ATen ops (pinned torch==2.5.1): stft, conv1d, layer_norm, gelu(erf), irfft, fold.
    The port calls the same ATen ops so that on one machine it is bit-identical to the reference
    in fp32; `dtype=torch.float64` gives a higher-precision truth for error budgeting.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

N_FFT = 400
HOP = 160
N_MELS = 80
N_SAMPLES = 480000
N_FRAMES = 3000


# ----------------------------------------------------------------------------------------------
# log-mel front end  (reference audiocodec/nn/feature_extractor.py:50-58, 86-112, 136-245)
# ----------------------------------------------------------------------------------------------

def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, lin)


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    lin = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), lin)


def mel_filterbank(n_bins: int = 201, n_mels: int = N_MELS, fmin: float = 0.0, fmax: float = 8000.0,
                   sr: int = 16000) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular filterbank, shape (n_bins, n_mels), float64.
    Restates transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") as called at
    feature_extractor.py:50-58."""
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0.0, sr // 2, n_bins)
    fdiff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[None, :]


def log_mel(wavs: Sequence[torch.Tensor], dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """list of 1-D waveforms (<= 480000 samples) -> (B,80,3000) features, (B,) mel lengths.
    feature_extractor.py:207-214 (pad to n_samples), :97-109 (stft/power/mel/log/clamp/affine),
    :236-237 + model.py:191 (mel_len = number of samples at multiples of the hop = ceil(L/160))."""
    B = len(wavs)
    x = torch.zeros(B, N_SAMPLES, dtype=dtype)
    lens = torch.zeros(B, dtype=torch.long)
    for i, w in enumerate(wavs):
        w = w.reshape(-1)[:N_SAMPLES]
        x[i, : w.numel()] = w.to(dtype)
        lens[i] = (w.numel() + HOP - 1) // HOP
    window = torch.hann_window(N_FFT, dtype=dtype)
    stft = torch.stft(x, N_FFT, HOP, window=window, return_complex=True)
    power = stft[..., :-1].abs() ** 2
    fb = torch.from_numpy(mel_filterbank()).to(torch.float32).to(dtype)   # fp64 -> fp32 cast at :100
    mel = fb.T @ power
    logm = torch.clamp(mel, min=1e-10).log10()
    mx = logm.max(dim=2, keepdim=True)[0].max(dim=1, keepdim=True)[0]
    logm = torch.maximum(logm, mx - 8.0)
    return (logm + 4.0) / 4.0, lens


# ----------------------------------------------------------------------------------------------
# transformer blocks  (reference audiocodec/nn/modules.py:85-232)
# ----------------------------------------------------------------------------------------------

def _lin(sd: SD, p: str, x: torch.Tensor, bias: bool = True) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"] if bias else None)


def _ln(sd: SD, p: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def attention(sd: SD, p: str, h: torch.Tensor, lens: torch.Tensor, heads: int) -> torch.Tensor:
    """VarLenAttention.forward (modules.py:145-187) with the additive mask of :111-143
    (+1 on valid pairs, finfo.min elsewhere)."""
    B, T, D = h.shape
    hd = D // heads
    q = _lin(sd, p + ".q_proj", h) * hd ** -0.5
    k = _lin(sd, p + ".k_proj", h, bias=False)
    v = _lin(sd, p + ".v_proj", h)
    q, k, v = (t.view(B, T, heads, hd).transpose(1, 2) for t in (q, k, v))
    scores = torch.matmul(q, k.transpose(-1, -2))
    valid = torch.arange(T)[None, :] < lens[:, None]                       # (B,T)
    pair = (valid[:, None, :, None] & valid[:, None, None, :]).to(h.dtype)  # (B,1,T,T)
    mask = pair + (1.0 - pair) * torch.finfo(h.dtype).min
    w = F.softmax(scores + mask, dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, T, D)
    return _lin(sd, p + ".out_proj", o)


def transformer_layer(sd: SD, p: str, h: torch.Tensor, lens: torch.Tensor, heads: int) -> torch.Tensor:
    """Pre-LN Whisper layer (modules.py:214-232); LayerNorm eps 1e-5, exact-erf GELU."""
    h = h + attention(sd, p + ".self_attn", _ln(sd, p + ".self_attn_layer_norm", h, 1e-5), lens, heads)
    m = _ln(sd, p + ".final_layer_norm", h, 1e-5)
    return h + _lin(sd, p + ".fc2", F.gelu(_lin(sd, p + ".fc1", m)))


def _n_layers(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.layers.{n}.fc1.weight" in sd:
        n += 1
    return n


def encoder(sd: SD, mel: torch.Tensor, mel_lens: torch.Tensor, heads: int = 12,
            return_stem: bool = False, output_hidden_states: bool = False):
    """OmniAudioEncoder.forward with is_acoustic=True (modules.py:287-376): two un-activated convs,
    no positional embedding, N layers, final LN, zero the rows >= len, channels-first output.
    output_hidden_states (modules.py:344-371): also the masked, transposed input of every layer and the final LN output."""
    p = "acoustic_encoder"
    x = F.conv1d(mel, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    x = F.conv1d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], stride=2, padding=1)
    lens = (mel_lens // 2).long()
    h = x.permute(0, 2, 1)
    stem = h
    hidden = []
    for i in range(_n_layers(sd, p)):
        hidden.append(h)
        h = transformer_layer(sd, f"{p}.layers.{i}", h, lens, heads)
    h = _ln(sd, p + ".layer_norm", h, 1e-5)
    hidden.append(h)
    keep = (torch.arange(h.shape[1])[None, :] < lens[:, None])[..., None]
    mask = lambda t: torch.where(keep, t, torch.zeros((), dtype=t.dtype)).transpose(1, 2)  # noqa: E731
    h = mask(h)
    if output_hidden_states:
        return h, lens, tuple(mask(t) for t in hidden)
    return (h, lens, stem) if return_stem else (h, lens)


# ----------------------------------------------------------------------------------------------
# frame-stack resamplers  (modules.py:37-49, 476-634; activations.py:107-119; alias_free_torch/*)
# ----------------------------------------------------------------------------------------------

def _wn_weight(sd: SD, p: str) -> torch.Tensor:
    """Old-style weight norm, norm over (in, k) per out-channel: w = g * v / ||v||."""
    v, g = sd[p + ".weight_v"], sd[p + ".weight_g"]
    return v * (g / torch.linalg.vector_norm(v, dim=(1, 2), keepdim=True))


def _wn_conv(sd: SD, p: str, x: torch.Tensor, dilation: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv1d(x, _wn_weight(sd, p), sd[p + ".bias"], dilation=dilation, padding=padding)


def aa_snake(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """Activation1d(SnakeBeta) (act.py:23-27): 2x Kaiser-sinc upsample (resample.py:25-33),
    x + sin^2(x e^a)/(e^b + 1e-9) (activations.py:107-119), 2x low-pass downsample (filter.py:83-92)."""
    C = x.shape[1]
    fu = sd[p + ".upsample.filter"].to(x.dtype).expand(C, -1, -1)
    fd = sd[p + ".downsample.lowpass.filter"].to(x.dtype).expand(C, -1, -1)
    u = F.pad(x, (5, 5), mode="replicate")
    u = 2 * F.conv_transpose1d(u, fu, stride=2, groups=C)
    u = u[..., 15:-15]
    a = torch.exp(sd[p + ".act.alpha"])[None, :, None]
    b = torch.exp(sd[p + ".act.beta"])[None, :, None]
    u = u + (1.0 / (b + 1e-9)) * torch.sin(u * a) ** 2
    u = F.pad(u, (5, 6), mode="replicate")
    return F.conv1d(u, fd, stride=2, groups=C)


def residual_unit(sd: SD, p: str, x: torch.Tensor, dilation: int) -> torch.Tensor:
    y = aa_snake(sd, p + ".block.0", x)
    y = _wn_conv(sd, p + ".block.1", y, dilation=dilation, padding=3 * dilation)
    y = aa_snake(sd, p + ".block.2", y)
    return x + _wn_conv(sd, p + ".block.3", y)


def downsample(sd: SD, x: torch.Tensor, lens: torch.Tensor, s: int = 4) -> Tuple[torch.Tensor, torch.Tensor]:
    """FrameStackDownConv.forward (modules.py:519-550). x (B,D,T) -> (B,32,ceil(T/s))."""
    B, D, T = x.shape
    out_len = (lens + s - 1) // s
    Tp = (T + s - 1) // s * s
    if Tp > T:
        x = F.pad(x, (0, Tp - T))
    x = x.view(B, D, Tp // s, s).permute(0, 1, 3, 2).reshape(B, D * s, Tp // s)   # ch = d*s + j
    h = _wn_conv(sd, "downsample.in_proj", x)
    for i, d in enumerate((1, 3, 9)):
        h = residual_unit(sd, f"downsample.res_blocks.{i}", h, d)
    return _wn_conv(sd, "downsample.to_latent", h), out_len


def upsample(sd: SD, zq: torch.Tensor, lens: torch.Tensor, s: int = 4) -> Tuple[torch.Tensor, torch.Tensor]:
    """FrameStackUpConv.forward (modules.py:601-631). (B,32,T') -> (B,768,4T')."""
    h = _wn_conv(sd, "upsample.from_latent", zq)
    for i, d in enumerate((1, 3, 9)):
        h = residual_unit(sd, f"upsample.res_blocks.{i}", h, d)
    h = _wn_conv(sd, "upsample.to_stacked", h)
    B, DS, T = h.shape
    y = h.view(B, DS // s, s, T).permute(0, 1, 3, 2).reshape(B, DS // s, T * s)
    return y, lens * s


# ----------------------------------------------------------------------------------------------
# group FSQ  (reference audiocodec/nn/quantizer.py:9-30, 121-224, 273-317)
# ----------------------------------------------------------------------------------------------

FSQ_LEVELS = (8, 7, 6, 6)
FSQ_EPS = 1e-3


def _len_mask(T: int, lens: torch.Tensor) -> torch.Tensor:
    return torch.arange(1, T + 1)[None, None, :] <= lens[:, None, None]


def fsq_encode(z: torch.Tensor, lens: torch.Tensor, levels=FSQ_LEVELS, eps: float = FSQ_EPS):
    """z (B, G*4, T) -> dequantised (B, G*4, T), indices (G, B, T) int32; positions >= len zeroed.
    Per group: compress 129-140, round-half-even 121-127, normalise 155-156, index 159-179."""
    B, C, T = z.shape
    D = len(levels)
    G = C // D
    L = torch.tensor(levels, dtype=torch.int32).view(1, D, 1)
    base = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0).to(torch.int32).view(1, D, 1)
    keep = _len_mask(T, lens)
    dq_all, idx_all = [], []
    for g in range(G):
        x = z[:, g * D:(g + 1) * D]
        scale = (L - 1) / 2 * (1 - eps)
        offset = torch.where(L % 2 == 0, 0.5, 0)
        shift = (offset / scale).tan()
        c = scale * (x + shift).tanh() - offset
        r = torch.round(c)
        half = L // 2
        dq = r / half
        idx = torch.sum((half * dq + half) * base, dim=1).to(torch.int32)
        dq_all.append(dq * keep)
        idx_all.append((idx * keep[:, 0]).unsqueeze(0))
    return torch.cat(dq_all, dim=1), torch.cat(idx_all, dim=0)


def fsq_decode(idx: torch.Tensor, lens: torch.Tensor, levels=FSQ_LEVELS, dtype=torch.float32) -> torch.Tensor:
    """indices (G,B,T) -> (B, G*4, T)  (quantizer.py:207-224, 306-317)."""
    G, B, T = idx.shape
    D = len(levels)
    L = torch.tensor(levels, dtype=torch.int32).view(1, D, 1)
    base = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0).to(torch.int32).view(1, D, 1)
    keep = _len_mask(T, lens)
    outs = []
    for g in range(G):
        nn_ = (idx[g][:, None, :] // base) % L
        half = L // 2
        dq = (nn_ - half) / half
        outs.append(dq * keep)
    return torch.cat(outs, dim=1).to(dtype)


# ----------------------------------------------------------------------------------------------
# decoder + Vocos + iSTFT  (modules.py:437-474, 1229-1248, 1492-1504, 1064-1082, 831-886)
# ----------------------------------------------------------------------------------------------

def decoder(sd: SD, x: torch.Tensor, lens: torch.Tensor, heads: int = 12) -> Tuple[torch.Tensor, torch.Tensor]:
    """OmniAudioDecoder.forward: (B,768,T) -> (B,80,2T)."""
    p = "acoustic_decoder"
    h = x.transpose(1, 2)
    T = h.shape[1]
    for i in range(_n_layers(sd, p)):
        h = transformer_layer(sd, f"{p}.layers.{i}", h, lens, heads)
    h = _ln(sd, p + ".layer_norm", h, 1e-5)
    keep = (torch.arange(T)[None, :] < lens[:, None])[..., None]
    h = torch.where(keep, h, torch.zeros((), dtype=h.dtype)).permute(0, 2, 1)
    y = F.conv_transpose1d(h, sd[p + ".deconv1.weight"], sd[p + ".deconv1.bias"], stride=2)
    y = F.conv_transpose1d(y, sd[p + ".deconv2.weight"], sd[p + ".deconv2.bias"], stride=1)
    return y[:, :, : 2 * T], lens * 2


def vocos_backbone(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """VocosBackbone.forward: (B,80,T) -> (B,T,512); no length masking anywhere."""
    p = "vocos.backbone"
    x = F.conv1d(x, sd[p + ".embed.weight"], sd[p + ".embed.bias"], padding=3)
    x = _ln(sd, p + ".norm", x.transpose(1, 2), 1e-6).transpose(1, 2)
    i = 0
    while f"{p}.convnext.{i}.gamma" in sd:
        q = f"{p}.convnext.{i}"
        C = x.shape[1]
        y = F.conv1d(x, sd[q + ".dwconv.weight"], sd[q + ".dwconv.bias"], padding=3, groups=C)
        y = _ln(sd, q + ".norm", y.transpose(1, 2), 1e-6)
        y = _lin(sd, q + ".pwconv2", F.gelu(_lin(sd, q + ".pwconv1", y)))
        x = x + (sd[q + ".gamma"] * y).transpose(1, 2)
        i += 1
    return _ln(sd, p + ".final_layer_norm", x.transpose(1, 2), 1e-6)


def istft_head(sd: SD, h: torch.Tensor, n_fft: int = 640, hop: int = 160) -> torch.Tensor:
    """ISTFTHead.forward + ISTFT.forward("same"): (B,T,512) -> (B, hop*T)."""
    x = _lin(sd, "vocos.head.out", h).transpose(1, 2)
    mag, ph = x.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = mag * (torch.cos(ph) + 1j * torch.sin(ph))
    win = sd["vocos.head.istft.window"].to(h.dtype)
    B, N, T = S.shape
    pad = (n_fft - hop) // 2
    fr = torch.fft.irfft(S, n_fft, dim=1, norm="backward") * win[None, :, None]
    out_size = (T - 1) * hop + n_fft
    y = F.fold(fr, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, pad:-pad]
    wsq = win.square().expand(1, T, -1).transpose(1, 2)
    env = F.fold(wsq, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop)).squeeze()[pad:-pad]
    return y / env


def vocos(sd: SD, x: torch.Tensor, lens: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return istft_head(sd, vocos_backbone(sd, x))[:, None, :], lens * 160


# ----------------------------------------------------------------------------------------------
# codec API  (reference audiocodec/model.py:112-373)
# ----------------------------------------------------------------------------------------------

def cast_sd(sd: SD, dtype) -> SD:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def tokenize(sd: SD, x: torch.Tensor, lengths: torch.Tensor, dtype=torch.float32, trace: dict = None):
    """inference_tokenize (model.py:167-210): x (B,1,T<=480000) -> zq (B,32,375), codes (8,B,375), lens."""
    wavs = [xi[0, : int(n)] for xi, n in zip(x, lengths)]
    mel, mel_lens = log_mel(wavs, dtype)
    enc, enc_lens = encoder(sd, mel, mel_lens)
    lat, lat_lens = downsample(sd, enc, enc_lens)
    zq, codes = fsq_encode(lat, lat_lens)
    if trace is not None:
        trace.update(mel=mel, mel_lens=mel_lens, enc=enc, enc_lens=enc_lens, latent=lat)
    return {"zq": zq, "codes": codes, "codes_lengths": lat_lens}


def detokenize(sd: SD, codes: torch.Tensor, lens: torch.Tensor, dtype=torch.float32, trace: dict = None):
    """inference_detokenize (model.py:212-242): codes (8,B,T') -> y (B,1,1280 T')."""
    zq = fsq_decode(codes, lens, dtype=dtype)
    up, up_lens = upsample(sd, zq, lens)
    dec, dec_lens = decoder(sd, up, up_lens)
    y, out_lens = vocos(sd, dec, dec_lens)
    if trace is not None:
        trace.update(zq=zq, up=up, dec=dec)
    return {"y": y, "output_length": out_lens}


def forward_mel(sd: SD, mel: torch.Tensor, mel_lens: torch.Tensor):
    """AudioCodec.forward (model.py:112-165): mel features in, reconstructed audio out, one pass."""
    enc, enc_lens = encoder(sd, mel, mel_lens)
    lat, lat_lens = downsample(sd, enc, enc_lens)
    zq, codes = fsq_encode(lat, lat_lens)
    up, up_lens = upsample(sd, zq, lat_lens)
    dec, dec_lens = decoder(sd, up, up_lens)
    y, out_lens = vocos(sd, dec, dec_lens)
    return {"reconstructed_audio": y, "audio_lengths": out_lens, "codes": codes}


def encode(sd: SD, wav_list: List[torch.Tensor], overlap_seconds: int = 10, sr: int = 16000,
           rate: int = 1280, max_seconds: int = 30, dtype=torch.float32) -> List[torch.Tensor]:
    """AudioCodec.encode (model.py:244-308): 30 s windows every (30-overlap) s, keep-first stitching."""
    keep_samples = (max_seconds - overlap_seconds) * sr
    win = max_seconds * sr
    keep_codes = keep_samples // rate
    B = len(wav_list)
    L = torch.tensor([len(w) for w in wav_list], dtype=torch.long)
    maxlen = int(L.max())
    x = torch.zeros(B, 1, maxlen, dtype=dtype)
    for i, w in enumerate(wav_list):
        x[i, 0, : len(w)] = w.to(dtype)
    pieces = []
    for c in range((maxlen + keep_samples - 1) // keep_samples):
        s, e = c * keep_samples, min(c * keep_samples + win, maxlen)
        cl = torch.clamp(L - s, 0, e - s)
        if int(cl.max()) == 0:
            continue
        r = tokenize(sd, x[:, :, s:e], cl, dtype)
        vl = torch.clamp(r["codes_lengths"], 0, keep_codes)
        piece = torch.zeros(r["codes"].shape[0], B, keep_codes, dtype=r["codes"].dtype)
        for b in range(B):
            piece[:, b, : int(vl[b])] = r["codes"][:, b, : int(vl[b])]
        pieces.append(piece)
    if not pieces:
        return [torch.zeros(8, 0, dtype=torch.long) for _ in range(B)]
    allc = torch.cat(pieces, dim=-1)
    return [allc[:, i, : int(L[i]) // rate] for i in range(B)]


def decode(sd: SD, codes_list: List[torch.Tensor], overlap_seconds: int = 10, sr: int = 16000,
           rate: int = 1280, max_seconds: int = 30, dtype=torch.float32) -> List[torch.Tensor]:
    """AudioCodec.decode (model.py:310-373): windows of <=375 codes every 250, NOT padded to 375 —
    the window length T' is the batch maximum, and the un-masked convs see the zero padding."""
    win_codes = max_seconds * sr // rate
    keep_codes = (max_seconds - overlap_seconds) * sr // rate
    keep_wav = keep_codes * rate
    B = len(codes_list)
    G = codes_list[0].shape[0]
    L = torch.tensor([c.shape[-1] for c in codes_list], dtype=torch.long)
    maxlen = int(L.max())
    ct = torch.zeros(G, B, maxlen, dtype=torch.long)
    for i, c in enumerate(codes_list):
        ct[:, i, : c.shape[-1]] = c
    pieces = []
    for c in range((maxlen + keep_codes - 1) // keep_codes):
        s, e = c * keep_codes, min(c * keep_codes + win_codes, maxlen)
        cl = torch.clamp(L - s, 0, e - s)
        if int(cl.max()) == 0:
            continue
        r = detokenize(sd, ct[:, :, s:e], cl, dtype)
        vl = torch.clamp(r["output_length"], 0, keep_wav)
        piece = torch.zeros(B, 1, keep_wav, dtype=dtype)
        for b in range(B):
            piece[b, :, : int(vl[b])] = r["y"][b, :, : int(vl[b])]
        pieces.append(piece)
    if not pieces:
        return [torch.zeros(0, dtype=dtype) for _ in range(B)]
    allw = torch.cat(pieces, dim=-1)
    return [allw[i, 0, : int(L[i]) * rate] for i in range(B)]

"""Drop-in `AudioCodec` for the reference's `audiocodec/model.py` (ZhangXinWhut/SimWhisper-Codec),
backed by the sm_100a kernels of libswc.so.

Same constructor (`AudioCodec(generator_params)`), attributes, sub-module names, `state_dict()` key
schema and inference API as the reference (audiocodec/model.py:15-396):
`encode / decode / inference_tokenize / inference_detokenize / forward / load_from_checkpoint`.
The sub-modules hold the parameters and call the C ABI through `NativeCodec`; there is no PyTorch or
CPU compute path — without a CUDA device (or without the built library) every forward raises.

Differences that are deliberate and documented in DESIGN.md:
  * `encode()/decode()` flatten all (item x 30-s window) pairs into one batch instead of looping
    over windows in Python, and never synchronise with the host (lengths are known on the host);
    results are identical to the reference's windowing (keep-first 20 s, decode pad length T' =
    batch maximum per window index, reference model.py:275-297, 340-362).
  * `precision="bf16"` selects the tcgen05 tensor-core path (fp32 residual stream, LayerNorm,
    softmax, iSTFT); `precision="fp32"` is the parity mode (CUDA-core FFMA everywhere);
    `precision="bf16x3"` is the parity mode with the dense contractions on the tensor cores as three
    bf16 products with fp32 accumulation (fp32 activations, attention, mel and iSTFT as in "fp32").
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import yaml

from .. import _lib, windows
from ..weights import state_dict_schema

# encode()/decode() know every window's length on the host: skip padded tokens in the transformer stacks (bf16 mode).
_RAGGED = os.environ.get("SWC_RAGGED", "1") != "0"
_BUFFER_LEAVES = ("positional_embedding", "filter", "window", "dim_base_index", "num_levels")


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(device=None) -> C.c_void_p:
    """The caller's current stream ON THE TENSORS' DEVICE (not on whatever device happens to be current)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class NativeCodec:
    """Owns the swc_model handle and the scratch workspace of one device."""

    def __init__(self, precision: str):
        self.lib = _lib.load()
        self.precision = precision
        self.handle = C.c_void_p()
        _lib.check(self.lib.swc_model_create(C.byref(self.handle), _lib.PRECISION[precision]), "swc_model_create")
        self.device: Optional[torch.device] = None
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.swc_model_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:  # pragma: no cover - interpreter teardown
            pass

    def set_state(self, sd: Dict[str, torch.Tensor]) -> None:
        for k, v in sd.items():
            t = v.detach().to("cpu").contiguous()
            if t.dtype in (torch.int32,):
                dt = 1
            else:
                t = t.to(torch.float32)
                dt = 0
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            _lib.check(self.lib.swc_model_set_tensor(self.handle, k.encode(), _ptr(t), dt, shape, t.dim()),
                       f"swc_model_set_tensor({k})")

    def pack(self) -> None:
        _lib.check(self.lib.swc_model_pack(self.handle), "swc_model_pack")

    def packed(self, name: str) -> torch.Tensor:
        n = self.lib.swc_model_packed_numel(self.handle, name.encode())
        if n < 0:
            raise KeyError(name)
        out = torch.empty(n, dtype=torch.float32)
        _lib.check(self.lib.swc_model_get_packed(self.handle, name.encode(), _ptr(out), n), "swc_model_get_packed")
        return out

    def finalize(self, device: torch.device) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("SimWhisper-Codec B200 path needs a CUDA device; there is no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(self.lib.swc_model_finalize(self.handle, idx), "swc_model_finalize")
        self.device = torch.device("cuda", idx)

    def workspace(self, stage: str, batch: int, frames: int):
        need = int(self.lib.swc_workspace_bytes(self.handle, _lib.STAGE[stage], batch, frames))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return _ptr(self._ws), C.c_size_t(self._ws.numel())


class _GraphBucket:
    """One captured call of the padded single-pass chain for a fixed (stage, batch, frames) bucket: static input / output /
    workspace buffers, replayed with a single launch.  At one or a few windows a call is otherwise launch-bound on the host
    (~880 launches per tokenize + detokenize at ~5 us each); the library allocates nothing, never synchronises and reads
    nothing back, so the whole call captures as it is."""

    def __init__(self, run, inputs, outputs, ws):
        self.inputs, self.outputs, self.ws = inputs, outputs, ws
        self.dev = dev = ws.device
        # capture and replay with the buffers' device current: torch captures on (and replays into) a stream of the CURRENT
        # device, and a model on cuda:1 called while cuda:0 is current would otherwise capture an empty graph
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                run()                                        # warm-up outside the capture (function attributes, caches)
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            # (an explicit capture stream: torch.cuda.graph otherwise reuses one process-wide default stream, which lives
            # on the device of the FIRST capture)
            with torch.cuda.graph(self.graph, stream=side):
                run()

    def replay(self):
        with torch.cuda.device(self.dev):
            self.graph.replay()


class _Holder(nn.Module):
    """Plain container that reproduces the reference module tree so state_dict keys line up."""


class _Stage(_Holder):
    def __init__(self, owner: "AudioCodec", name: str):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        object.__setattr__(self, "_stage", name)

    def _run(self, fn_name: str, stage: str, x: torch.Tensor, lens: torch.Tensor, out_shape, frames: int):
        owner = self._owner
        nat = owner._native_for(x.device)
        x = x.contiguous().to(torch.float32)
        lens = lens.to(device=x.device, dtype=torch.int64).contiguous()
        B = x.shape[0]
        out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
        out_lens = torch.empty(B, dtype=torch.int64, device=x.device)
        ws, nws = nat.workspace(stage, B, frames)
        fn = getattr(nat.lib, fn_name)
        _lib.check(fn(nat.handle, _ptr(x), _ptr(lens), B, frames, _ptr(out), _ptr(out_lens), ws, nws, _stream(x.device)), fn_name)
        return out, out_lens


class OmniAudioEncoder(_Stage):
    """reference audiocodec/nn/modules.py:236-376 — forward(input_features (B,80,T), input_length)."""

    def forward(self, input_features, input_length, output_hidden_states=False):
        B, _, T = input_features.shape
        if not output_hidden_states:
            return self._run("swc_encoder", "encoder", input_features, input_length, (B, 768, (T + 1) // 2), T)
        # reference modules.py:344-371: additionally the input of every layer and the final LayerNorm output, masked
        nat = self._owner._native_for(input_features.device)
        x = input_features.contiguous().to(torch.float32)
        lens = input_length.to(device=x.device, dtype=torch.int64).contiguous()
        n_layers = sum(1 for k in self._owner.state_dict() if k.startswith("acoustic_encoder.layers.") and k.endswith(".fc1.bias"))
        out = torch.empty((B, 768, (T + 1) // 2), dtype=torch.float32, device=x.device)
        hidden = torch.empty((n_layers + 1, B, 768, (T + 1) // 2), dtype=torch.float32, device=x.device)
        out_lens = torch.empty(B, dtype=torch.int64, device=x.device)
        ws, nws = nat.workspace("encoder", B, T)
        _lib.check(nat.lib.swc_encoder_hidden(nat.handle, _ptr(x), _ptr(lens), B, T, _ptr(out), _ptr(out_lens), _ptr(hidden),
                                              ws, nws, _stream(x.device)), "swc_encoder_hidden")
        return out, out_lens, tuple(hidden.unbind(0))


class FrameStackDownConv(_Stage):
    """reference modules.py:476-553 — forward(x (B,768,T), input_length) -> (B,32,ceil(T/4))."""

    def forward(self, x, input_length):
        B, _, T = x.shape
        return self._run("swc_downsample", "downsample", x, input_length, (B, 32, (T + 3) // 4), T)


class FrameStackUpConv(_Stage):
    """reference modules.py:555-634 — forward(z_q (B,32,T'), input_len) -> (B,768,4T')."""

    def forward(self, z_q, input_len=None):
        B, _, T = z_q.shape
        if input_len is None:
            input_len = torch.full((B,), T, dtype=torch.int64, device=z_q.device)
        return self._run("swc_upsample", "upsample", z_q, input_len, (B, 768, 4 * T), T)


class OmniAudioDecoder(_Stage):
    """reference modules.py:380-474 — forward(hidden_states (B,768,T), input_length) -> (B,80,2T)."""

    def forward(self, hidden_states, input_length):
        B, _, T = hidden_states.shape
        return self._run("swc_decoder", "decoder", hidden_states, input_length, (B, 80, 2 * T), T)


class Vocos(_Stage):
    """reference modules.py:1545-1573 — forward(x (B,80,T), input_length) -> (B,1,160T)."""

    def forward(self, x, input_length):
        B, _, T = x.shape
        y, lens = self._run("swc_vocos", "vocos", x, input_length, (B, 160 * T), T)
        return y[:, None, :], lens


class GroupFiniteScalarQuantizer(_Stage):
    """reference audiocodec/nn/quantizer.py:226-317."""

    def forward(self, inputs, input_len):
        if inputs.size(1) != 32:
            raise RuntimeError(
                f"Input dimension {inputs.size(1)} not matching the expected dimension 32, inputs shape {inputs.shape}")
        nat = self._owner._native_for(inputs.device)
        x = inputs.contiguous().to(torch.float32)
        lens = input_len.to(device=x.device, dtype=torch.int64).contiguous()
        B, _, T = x.shape
        zq = torch.empty_like(x)
        codes = torch.empty((8, B, T), dtype=torch.int32, device=x.device)
        _lib.check(nat.lib.swc_quantize(nat.handle, _ptr(x), _ptr(lens), B, T, _ptr(zq), _ptr(codes), _stream(x.device)), "swc_quantize")
        return zq, codes

    def encode(self, inputs, input_len):
        return self.forward(inputs, input_len)[1]

    def decode(self, indices, input_len):
        if indices.dim() != 3 or indices.size(0) != 8:
            raise ValueError(f"Expected indices of shape (8, B, T), got {tuple(indices.shape)}")
        nat = self._owner._native_for(indices.device)
        if indices.dtype not in (torch.int32, torch.int64):
            indices = indices.to(torch.int64)
        idx = indices.contiguous()
        lens = input_len.to(device=idx.device, dtype=torch.int64).contiguous()
        _, B, T = idx.shape
        zq = torch.empty((B, 32, T), dtype=torch.float32, device=idx.device)
        _lib.check(nat.lib.swc_dequantize(nat.handle, _ptr(idx), int(idx.dtype == torch.int64), _ptr(lens), B, T,
                                          _ptr(zq), _stream(idx.device)), "swc_dequantize")
        return zq


class MelFeatureExtractor:
    """reference audiocodec/nn/feature_extractor.py:19-245, computed on the GPU.  Accepts the same
    list-of-1-D-arrays input and returns {"input_features" (B,80,3000), "attention_mask" (B,3000)}."""

    def __init__(self, owner: "AudioCodec", **kwargs):
        self._owner = owner
        self.n_samples = 480000
        self.nb_max_frames = 3000
        self.hop_length = 160

    def __call__(self, raw_speech, sampling_rate=None, return_tensors="pt", return_attention_mask=True,
                 device="cuda", **_):
        dev = torch.device(device if device != "cpu" else "cuda")
        wavs = [torch.as_tensor(w, dtype=torch.float32).reshape(-1)[: self.n_samples] for w in raw_speech]
        B = len(wavs)
        L = max(1, max(w.numel() for w in wavs))
        x = torch.zeros(B, L, dtype=torch.float32, device=dev)
        for i, w in enumerate(wavs):
            x[i, : w.numel()] = w.to(dev)
        lens = torch.tensor([w.numel() for w in wavs], dtype=torch.int64, device=dev)
        mel, mel_lens = self._owner._mel(x, lens)
        out = {"input_features": mel}
        if return_attention_mask:
            out["attention_mask"] = (torch.arange(3000, device=dev)[None, :] < mel_lens[:, None]).to(torch.int32)
        return out


class AudioCodec(nn.Module):
    def __init__(self, generator_params: dict, precision: str = None, max_batch: int = None):
        super().__init__()
        gp = generator_params
        self.input_sample_rate = gp["input_sample_rate"]
        self.output_sample_rate = gp["output_sample_rate"]
        self.max_audio_seconds = 30
        self.encoder_downsample_rate = gp["encoder_downsample_rate"]
        self.decoder_upsample_rate = gp["decoder_upsample_rate"]
        self.num_groups = gp["quantizer"]["num_groups"]
        self.codebook_dim_per_group = len(gp["quantizer"]["num_levels_per_group"])
        self._validate(gp)
        # reference model.py:36-39: Whisper initialisation options ride in the encoder's config block
        self.freeze_acoustic_encoder_flag = gp["acoustic_encoder"].get("freeze", False)
        self.whisper_model_path = gp["acoustic_encoder"].get("whisper_model_path", None)
        self.init_from_whisper = gp["acoustic_encoder"].get("init_from_whisper", False)
        # default: the parity-grade tensor-core mode (meets the fp32 bars of north_star at tensor-core speed); "bf16" is the
        # throughput mode, "fp32" the CUDA-core reference arithmetic
        self.precision = precision or os.environ.get("SWC_PRECISION", "bf16x3")
        if self.precision not in _lib.PRECISION:
            raise ValueError(f"precision must be one of {list(_lib.PRECISION)}, got {self.precision}")
        self.max_batch = max(1, int(max_batch or os.environ.get("SWC_MAX_BATCH", 128)))

        self.acoustic_encoder = OmniAudioEncoder(self, "encoder")
        self.downsample = FrameStackDownConv(self, "downsample")
        self.quantizer = GroupFiniteScalarQuantizer(self, "quantizer")
        self.upsample = FrameStackUpConv(self, "upsample")
        self.acoustic_decoder = OmniAudioDecoder(self, "decoder")
        self.vocos = Vocos(self, "vocos")
        self.feature_extractor = MelFeatureExtractor(self, **gp["feature_extractor"])

        # parameter / buffer tree with the reference's key schema (zero-filled until load_state_dict)
        for key, (shape, dtype, _spec) in state_dict_schema(gp).items():
            parts = key.split(".")
            mod = self
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, _Holder())
                mod = getattr(mod, p)
            t = torch.zeros(shape, dtype=dtype)
            if parts[-1] in _BUFFER_LEAVES:
                mod.register_buffer(parts[-1], t)
            else:
                mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
        self._native: Optional[NativeCodec] = None
        self._native_version = -1
        self._version = 0
        # CUDA graph per bucket: calls of at most `graph_max_batch` windows replay a captured graph of the padded chain
        self.graph_max_batch = int(os.environ.get("SWC_GRAPH_MAX_BATCH", 4))
        self.api_chunk = int(os.environ.get("SWC_API_CHUNK", 96))      # windows per upload / launch chunk of encode()
        self._graphs: Dict[tuple, _GraphBucket] = {}
        self.graph_replays = 0
        self._copy_streams: Dict[int, "torch.cuda.Stream"] = {}      # per device: host -> device copies of encode()

    # ------------------------------------------------------------------ config / weights
    @staticmethod
    def _validate(gp: dict) -> None:
        enc, dec, dn, up, q, vo, fe = (gp[k] for k in ("acoustic_encoder", "acoustic_decoder", "downsample",
                                                       "upsample", "quantizer", "vocos", "feature_extractor"))
        ok = (enc["d_model"] == 768 and dec["d_model"] == 768 and enc["encoder_attention_heads"] == 12
              and dec["decoder_attention_heads"] == 12 and enc["encoder_ffn_dim"] == 3072
              and dec["decoder_ffn_dim"] == 3072 and enc["num_mel_bins"] == 80 and enc["stride_size"] == 2
              and enc["kernel_size"] == 3 and enc.get("is_acoustic", False)
              and dn["hidden_dim"] == 512 and up["hidden_dim"] == 512 and dn["stack_factor"] == 4
              and dn["latent_dim"] == 32 and q["num_groups"] == 8 and len(q["num_levels_per_group"]) == 4
              and abs(q.get("eps", 1e-3) - 1e-3) < 1e-12
              and vo["dim"] == 512 and vo["intermediate_dim"] == 4096 and vo["n_fft"] == 640
              and vo["hop_size"] == 160 and vo["padding"] == "same" and vo["input_channels"] == 80
              and fe["n_fft"] == 400 and fe["hop_length"] == 160 and fe["feature_size"] == 80
              and fe["sampling_rate"] == 16000 and gp["input_sample_rate"] == 16000)
        if not ok:
            raise ValueError("this build is specialised for config/SimWhisperCodec.yaml (768-d, 12 heads, 512-d "
                             "resamplers, 8x[8,7,6,6] FSQ, Vocos 512/4096, n_fft 640); other shapes have no kernels")

    def _init_whisper_weights(self):
        """reference model.py:59-88: copy the pretrained Whisper encoder into `acoustic_encoder` (utils/weight_init.py)."""
        if not self.init_from_whisper:
            return
        if self.whisper_model_path is None:
            logging.warning("init_from_whisper=True but no whisper_model_path given; skipping weight initialisation")
            return
        from ..utils.weight_init import load_whisper_weights
        load_whisper_weights(encoder=self.acoustic_encoder, whisper_model_name=self.whisper_model_path,
                             verbose=int(os.environ.get("RANK", 0)) == 0, is_acoustic=True)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        r = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._version += 1
        return r

    def _native_for(self, device: torch.device) -> NativeCodec:
        if device.type != "cuda":
            raise RuntimeError(f"SimWhisper-Codec B200 path runs on CUDA tensors only (got {device}); no CPU fallback")
        if self._native is None or self._native_version != self._version or self._native.device != torch.device(
                "cuda", device.index if device.index is not None else torch.cuda.current_device()):
            self._graphs.clear()                             # captured graphs hold pointers into the old weight slab
            self._graph_ws = None
            nat = NativeCodec(self.precision)
            nat.set_state(self.state_dict())
            nat.finalize(device)
            self._native, self._native_version = nat, self._version
        return self._native

    def _graph_workspace(self, nat, need: int, device) -> torch.Tensor:
        """One scratch buffer shared by all captured graphs of this model (they replay on one stream, one after the other),
        sized once for the largest bucket (graph_max_batch full windows); growing it would drop the graphs captured against
        the old buffer."""
        ws = getattr(self, "_graph_ws", None)
        if ws is None or ws.numel() < need or ws.device != device:
            n = max(1, self.graph_max_batch)
            need = max(need, int(nat.lib.swc_workspace_bytes(nat.handle, _lib.STAGE["tokenize"], n, 3000)),
                       int(nat.lib.swc_workspace_bytes(nat.handle, _lib.STAGE["detokenize"], n, 375)))
            self._graphs.clear()
            self._graph_ws = ws = torch.empty(need, dtype=torch.uint8, device=device)
        return ws

    def _bucket(self, key, build):
        b = self._graphs.get(key)
        if b is None:
            if len(self._graphs) >= 32:                      # bounded: drop the oldest bucket
                self._graphs.pop(next(iter(self._graphs)))
            b = self._graphs[key] = build()
        self.graph_replays += 1
        return b

    def pack_preview(self) -> NativeCodec:
        """Host-only packing (no GPU needed): used by CPU tests of the weight layout."""
        nat = NativeCodec(self.precision)
        nat.set_state(self.state_dict())
        nat.pack()
        return nat

    # ------------------------------------------------------------------ stage helpers
    def _mel(self, x2d: torch.Tensor, lens: torch.Tensor):
        nat = self._native_for(x2d.device)
        B = x2d.shape[0]
        mel = torch.empty((B, 80, 3000), dtype=torch.float32, device=x2d.device)
        mel_lens = torch.empty(B, dtype=torch.int64, device=x2d.device)
        ws, nws = nat.workspace("mel", B, 3000)
        _lib.check(nat.lib.swc_mel(nat.handle, _ptr(x2d), x2d.stride(0), x2d.shape[1], _ptr(lens), B, _ptr(mel),
                                   _ptr(mel_lens), ws, nws, _stream(x2d.device)), "swc_mel")
        return mel, mel_lens

    def _tokenize(self, x2d: torch.Tensor, lens: torch.Tensor, want_zq: bool, host_lens=None):
        """x2d (N, L<=480000) fp32 cuda, lens (N,) int64 cuda -> codes (8,N,375) int32, zq, code lens.
        host_lens (Python ints, when the caller knows them) lets the bf16 path skip the padded tokens of every window."""
        nat = self._native_for(x2d.device)
        N = x2d.shape[0]
        if 0 < N <= self.graph_max_batch and not torch.cuda.is_current_stream_capturing():
            return self._tokenize_graph(nat, x2d, lens, want_zq)
        codes = torch.empty((8, N, 375), dtype=torch.int32, device=x2d.device)
        zq = torch.empty((N, 32, 375), dtype=torch.float32, device=x2d.device) if want_zq else None
        clens = torch.empty(N, dtype=torch.int64, device=x2d.device)
        mb = self.max_batch
        if host_lens is not None and _RAGGED:
            mb = min(mb, int(nat.lib.swc_max_ragged()))      # the packed-token path takes at most this many items per call
        for s in range(0, N, mb):
            n = min(mb, N - s)
            ws, nws = nat.workspace("tokenize", n, 3000)
            c_part = codes if N <= mb else torch.empty((8, n, 375), dtype=torch.int32, device=x2d.device)
            z_part = None if zq is None else zq[s:s + n]
            if host_lens is not None and _RAGGED:
                hl = (C.c_int64 * n)(*[int(v) for v in host_lens[s:s + n]])
                _lib.check(nat.lib.swc_tokenize_ragged(nat.handle, _ptr(x2d[s:s + n]), x2d.stride(0), x2d.shape[1],
                                                       _ptr(lens[s:s + n]), hl, n, _ptr(c_part), _ptr(z_part),
                                                       _ptr(clens[s:s + n]), ws, nws, _stream(x2d.device)), "swc_tokenize_ragged")
            else:
                _lib.check(nat.lib.swc_tokenize(nat.handle, _ptr(x2d[s:s + n]), x2d.stride(0), x2d.shape[1], _ptr(lens[s:s + n]),
                                                n, _ptr(c_part), _ptr(z_part), _ptr(clens[s:s + n]), ws, nws, _stream(x2d.device)),
                           "swc_tokenize")
            if N > mb:
                codes[:, s:s + n] = c_part
        return codes, zq, clens

    def _tokenize_graph(self, nat, x2d, lens, want_zq):
        """small batches: the padded chain of this (batch, zq) bucket as one graph launch (results identical to the eager
        and to the packed-token paths, which are bit-equal to each other)."""
        dev, N, L = x2d.device, x2d.shape[0], min(x2d.shape[1], 480000)

        def build():
            x = torch.zeros((N, 480000), dtype=torch.float32, device=dev)
            ln = torch.zeros(N, dtype=torch.int64, device=dev)
            codes = torch.empty((8, N, 375), dtype=torch.int32, device=dev)
            zq = torch.empty((N, 32, 375), dtype=torch.float32, device=dev) if want_zq else None
            clens = torch.empty(N, dtype=torch.int64, device=dev)
            ws = self._graph_workspace(nat, int(nat.lib.swc_workspace_bytes(nat.handle, _lib.STAGE["tokenize"], N, 3000)), dev)

            def run():
                _lib.check(nat.lib.swc_tokenize(nat.handle, _ptr(x), x.stride(0), x.shape[1], _ptr(ln), N, _ptr(codes), _ptr(zq),
                                                _ptr(clens), _ptr(ws), C.c_size_t(ws.numel()), _stream(dev)), "swc_tokenize")
            return _GraphBucket(run, (x, ln), (codes, zq, clens), ws)

        b = self._bucket(("tokenize", dev.index, N, bool(want_zq)), build)
        x, ln = b.inputs
        x[:, :L].copy_(x2d[:, :L])                           # samples beyond an item's length are never read
        ln.copy_(lens.clamp(max=L))
        b.replay()
        codes, zq, clens = b.outputs
        return codes.clone(), (None if zq is None else zq.clone()), clens.clone()

    def _detokenize_graph(self, nat, codes, lens):
        dev = codes.device
        _, N, Tc = codes.shape
        i64 = codes.dtype == torch.int64

        def build():
            c = torch.zeros((8, N, Tc), dtype=codes.dtype, device=dev)
            ln = torch.zeros(N, dtype=torch.int64, device=dev)
            wav = torch.empty((N, 1280 * Tc), dtype=torch.float32, device=dev)
            olens = torch.empty(N, dtype=torch.int64, device=dev)
            ws = self._graph_workspace(nat, int(nat.lib.swc_workspace_bytes(nat.handle, _lib.STAGE["detokenize"], N, Tc)), dev)

            def run():
                _lib.check(nat.lib.swc_detokenize(nat.handle, _ptr(c), int(i64), _ptr(ln), N, Tc, _ptr(wav), _ptr(olens), _ptr(ws),
                                                  C.c_size_t(ws.numel()), _stream(dev)), "swc_detokenize")
            return _GraphBucket(run, (c, ln), (wav, olens), ws)

        b = self._bucket(("detokenize", dev.index, N, Tc, i64), build)
        c, ln = b.inputs
        c.copy_(codes)
        ln.copy_(lens)
        b.replay()
        wav, olens = b.outputs
        return wav.clone(), olens.clone()

    def _detokenize(self, codes: torch.Tensor, lens: torch.Tensor, host_lens=None):
        """codes (8,N,T') int32/int64 cuda -> wav (N, 1280 T'), out lens."""
        nat = self._native_for(codes.device)
        _, N, Tc = codes.shape
        if 0 < N <= self.graph_max_batch and Tc > 0 and not torch.cuda.is_current_stream_capturing():
            return self._detokenize_graph(nat, codes, lens)
        wav = torch.empty((N, 1280 * Tc), dtype=torch.float32, device=codes.device)
        olens = torch.empty(N, dtype=torch.int64, device=codes.device)
        mb = self.max_batch
        if host_lens is not None and _RAGGED:
            mb = min(mb, int(nat.lib.swc_max_ragged()))
        for s in range(0, N, mb):
            n = min(mb, N - s)
            ws, nws = nat.workspace("detokenize", n, Tc)
            c_part = codes if N <= mb else codes[:, s:s + n].contiguous()
            if host_lens is not None and _RAGGED:
                hl = (C.c_int64 * n)(*[int(v) for v in host_lens[s:s + n]])
                _lib.check(nat.lib.swc_detokenize_ragged(nat.handle, _ptr(c_part), int(codes.dtype == torch.int64),
                                                         _ptr(lens[s:s + n]), hl, n, Tc, _ptr(wav[s:s + n]),
                                                         _ptr(olens[s:s + n]), ws, nws, _stream(codes.device)), "swc_detokenize_ragged")
            else:
                _lib.check(nat.lib.swc_detokenize(nat.handle, _ptr(c_part), int(codes.dtype == torch.int64), _ptr(lens[s:s + n]), n,
                                                  Tc, _ptr(wav[s:s + n]), _ptr(olens[s:s + n]), ws, nws, _stream(codes.device)),
                           "swc_detokenize")
        return wav, olens

    # ------------------------------------------------------------------ reference API
    @torch.inference_mode()
    def forward(self, batch):
        """reference model.py:112-165 — {'mel_features' (B,80,T), 'mel_lens' (B,)} -> reconstructed audio."""
        mel = batch["mel_features"].contiguous().to(torch.float32)
        lens = batch["mel_lens"].to(device=mel.device, dtype=torch.int64).contiguous()
        nat = self._native_for(mel.device)
        B, _, Tm = mel.shape
        Tc = ((Tm + 1) // 2 + 3) // 4
        wav = torch.empty((B, 1280 * Tc), dtype=torch.float32, device=mel.device)
        olens = torch.empty(B, dtype=torch.int64, device=mel.device)
        ws, nws = nat.workspace("forward", B, Tm)
        _lib.check(nat.lib.swc_forward(nat.handle, _ptr(mel), _ptr(lens), B, Tm, _ptr(wav), _ptr(olens), _ptr(None),
                                       ws, nws, _stream(mel.device)), "swc_forward")
        return {"reconstructed_audio": wav[:, None, :], "audio_lengths": olens}

    @torch.inference_mode()
    def inference_tokenize(self, x, input_lengths):
        """reference model.py:167-210 — x (B,1,T<=30 s) -> zq (B,32,375), codes (8,B,375), codes_lengths."""
        x2d = x.reshape(x.shape[0], -1).contiguous().to(torch.float32)
        lens = input_lengths.to(device=x2d.device, dtype=torch.int64).contiguous()
        codes, zq, clens = self._tokenize(x2d, lens, want_zq=True)
        return {"zq": zq, "codes": codes, "codes_lengths": clens}

    @torch.inference_mode()
    def inference_detokenize(self, codes, codes_lengths):
        """reference model.py:212-242 — codes (8,B,T') -> y (B,1,1280 T'), output_length."""
        if codes.dtype not in (torch.int32, torch.int64):
            codes = codes.to(torch.int64)
        codes = codes.contiguous()
        lens = codes_lengths.to(device=codes.device, dtype=torch.int64).contiguous()
        wav, olens = self._detokenize(codes, lens)
        return {"y": wav[:, None, :], "output_length": olens}

    # ---- window batches (shared by the single-GPU API below and parallel.ShardedCodec)
    def _copy_stream(self, device) -> "torch.cuda.Stream":
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        st = self._copy_streams.get(idx)
        if st is None:
            st = self._copy_streams[idx] = torch.cuda.Stream(device=idx)
        return st

    def encode_jobs(self, wav_list, jobs, device) -> torch.Tensor:
        """Tokenize the given (item, start, n_valid) windows -> codes (8, len(jobs), 375) int32.

        Every window is copied host -> device exactly once, straight into its row of the chunk's input batch, on a copy
        stream of its own; the windows go through in chunks of at most `api_chunk`, each chunk's kernels wait only for
        that chunk's copies (an event), and nothing synchronises with the host.  From pinned memory the copies are
        asynchronous DMA and run under the kernels of the chunk before; from pageable memory the driver stages them
        through the host (the call then blocks while it copies, the GPU still overlaps).  A rank of a sharded job
        uploads only the windows it computes."""
        if not jobs:
            return torch.zeros((self.num_groups, 0, 375), dtype=torch.int32, device=device)
        device = torch.device(device)
        chunk = max(1, min(self.max_batch, self.api_chunk))
        cur = torch.cuda.current_stream(device)
        side = self._copy_stream(device)
        parts = [jobs[c0:c0 + chunk] for c0 in range(0, len(jobs), chunk)]
        # input batches first (allocated on the compute stream; the copy stream starts behind everything queued so far, so
        # a recycled block is never overwritten while an earlier kernel still reads it)
        xs = [torch.empty((len(p), max(j.n_valid for j in p)), dtype=torch.float32, device=device) for p in parts]
        flat = {}
        side.wait_stream(cur)
        outs = []
        for part, x in zip(parts, xs):
            with torch.cuda.stream(side):
                for k, j in enumerate(part):
                    src = flat.get(j.item)
                    if src is None:
                        src = flat[j.item] = torch.as_tensor(wav_list[j.item]).reshape(-1)
                    x[k, : j.n_valid].copy_(src[j.start:j.start + j.n_valid], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            cur.wait_event(ev)
            hl = [j.n_valid for j in part]
            wl = torch.tensor(hl, dtype=torch.int64).to(device, non_blocking=True)
            outs.append(self._tokenize(x, wl, want_zq=False, host_lens=hl)[0])      # samples beyond n_valid are never read
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)

    def decode_jobs(self, codes_list, jobs, device) -> torch.Tensor:
        """Detokenize decode windows that share one pad length T' -> wav (len(jobs), 1280 T')."""
        Tp = jobs[0].pad_len
        # the host knows every window's valid length: the library runs the transformer stack on the packed valid tokens and
        # Vocos (46 % of the FLOPs, no length masking in the reference) on each window's valid frames plus its
        # receptive-field halo, all windows packed into one batch of rows; samples beyond a window's halo are left
        # unwritten (callers keep the valid ones)
        ct = torch.zeros((self.num_groups, len(jobs), Tp), dtype=torch.int64, device=device)
        for k, j in enumerate(jobs):
            ct[:, k, : j.n_valid] = torch.as_tensor(codes_list[j.item])[:, j.start:j.start + j.n_valid].to(
                device=device, dtype=torch.int64, non_blocking=True)
        hl = [j.n_valid for j in jobs]
        cl = torch.tensor(hl, dtype=torch.int64).to(device, non_blocking=True)
        return self._detokenize(ct, cl, host_lens=hl)[0]

    @torch.inference_mode()
    def encode(self, wav_list, overlap_seconds=10, device=torch.device("cuda")):
        """reference model.py:244-308: 30-s windows every (30-overlap) s, first (30-overlap) s of codes kept.
        All (item, window) pairs run as one batch; stitching is a single gather (windows.py)."""
        device = torch.device(device)
        lens = [int(len(w)) for w in wav_list]
        jobs = windows.plan_encode(lens, overlap_seconds, self.input_sample_rate, self.max_audio_seconds)
        if not jobs:
            return {"codes_list": [torch.zeros(self.num_groups, 0, device=device, dtype=torch.long) for _ in lens]}
        codes = self.encode_jobs(wav_list, jobs, device)                       # (8, N, 375)
        return {"codes_list": self.stitch_codes(codes, lens, jobs, overlap_seconds)}

    def stitch_codes(self, codes, lens, jobs, overlap_seconds):
        src, splits = windows.encode_gather_index(lens, jobs, overlap_seconds, self.input_sample_rate,
                                                  self.max_audio_seconds, self.encoder_downsample_rate)
        flat = codes.reshape(self.num_groups, -1)
        if (src < 0).any():                  # positions the reference leaves zero (only when the hop is not a multiple of 1280)
            src = src.copy()
            src[src < 0] = flat.shape[1]
            flat = torch.cat([flat, flat.new_zeros(self.num_groups, 1)], dim=1)
        idx = torch.from_numpy(src).to(codes.device, non_blocking=True)
        return list(torch.split(flat.index_select(1, idx), splits, dim=1))

    @torch.inference_mode()
    def decode(self, codes_list, overlap_seconds=10, device=torch.device("cuda")):
        """reference model.py:310-373: windows of <=375 codes every 250; the pad length T' of window index c is the
        batch maximum min(375, maxlen-250c), exactly as the reference's un-padded decode batches."""
        device = torch.device(device)
        lens = [int(c.shape[-1]) for c in codes_list]
        up = self.decoder_upsample_rate
        # one flat buffer with a view per item (every sample is written by exactly one window), one fused copy for all windows
        outs = list(torch.split(torch.empty(sum(lens) * up, dtype=torch.float32, device=device), [L * up for L in lens]))
        groups = windows.plan_decode(lens, overlap_seconds, self.input_sample_rate, self.max_audio_seconds,
                                     self.encoder_downsample_rate)
        dst_views, src_views = [], []
        for _, jobs in groups.items():
            wav = self.decode_jobs(codes_list, jobs, device)
            for k, j in enumerate(jobs):
                off, n = windows.decode_keep(j, overlap_seconds, self.input_sample_rate, self.max_audio_seconds,
                                             self.encoder_downsample_rate, up)
                dst_views.append(outs[j.item][off:off + n])
                src_views.append(wav[k, :n])
        if dst_views:
            torch._foreach_copy_(dst_views, src_views)
        return {"syn_wav_list": outs}

    def remove_weight_norm(self):
        """API parity with the reference (modules.py remove_weight_norm on the resamplers): the packed weights already
        hold g * v / ||v|| folded once at load time, so there is nothing to remove."""
        return self

    @classmethod
    def load_from_checkpoint(cls, config_path: str, ckpt_path: str, **kwargs):
        """reference model.py:375-396 — YAML + checkpoint ({'model': sd} or bare state dict), strict."""
        logging.info(f"Loading model from {config_path} and {ckpt_path}")
        with open(config_path, "r") as f:
            config = yaml.safe_load(f)
        model = cls(config["generator_params"], **kwargs)
        checkpoint = torch.load(ckpt_path, map_location="cpu")
        model.load_state_dict(checkpoint["model"] if "model" in checkpoint else checkpoint, strict=True)
        return model

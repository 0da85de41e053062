"""Whisper-encoder weight initialisation for the acoustic encoder — mirror of the reference's
`utils/weight_init.py:11-76` (`load_whisper_weights`).

The reference copies every tensor of HF `WhisperModel(...).encoder.state_dict()` whose KEY also exists in the SimWhisper
encoder's state dict, skipping `positional_embedding` (both sides use the sinusoidal table, and the acoustic encoder never
adds it: modules.py:330-338).  That is 186 tensors: `conv1.{weight,bias}`, `conv2.{weight,bias}`, the 12 layers'
`self_attn.{q,v,out}_proj.{weight,bias}`, `self_attn.k_proj.weight`, `self_attn_layer_norm.*`, `fc1.*`, `fc2.*`,
`final_layer_norm.*` and the final `layer_norm.*`; only Whisper's `embed_positions.weight` has no counterpart
(tests/golden/whisper_init.json records the key set the reference's own function overwrites).

There is no network in this build environment, so besides a model name / local path (resolved through `transformers`
exactly as the reference does) the source may be given directly as a state dict of the Whisper ENCODER (keys as above,
with or without a leading `encoder.` / `model.encoder.` prefix).
"""
from __future__ import annotations

from typing import Dict, Mapping, Union

import torch

_PREFIXES = ("model.encoder.", "encoder.")


def whisper_encoder_state_dict(source: Union[str, Mapping[str, torch.Tensor]], local_files_only: bool = False) -> Dict[str, torch.Tensor]:
    """HF Whisper encoder state dict from a model name / local path, or a (possibly prefixed) mapping."""
    if isinstance(source, Mapping):
        out = {}
        for k, v in source.items():
            for p in _PREFIXES:
                if k.startswith(p):
                    k = k[len(p):]
                    break
            else:
                if k.startswith(("decoder.", "model.decoder.", "proj_out.")):
                    continue
            out[k] = v
        return out
    from transformers import WhisperModel          # reference utils/weight_init.py:8,33-50
    try:
        model = WhisperModel.from_pretrained(source, local_files_only=local_files_only)
    except Exception as e:
        if local_files_only:
            raise RuntimeError(f"Failed to load Whisper model from {source}: {e}")
        try:
            model = WhisperModel.from_pretrained(source, local_files_only=True)
        except Exception as e2:
            raise RuntimeError(f"Failed to load Whisper model from {source}: {e2}")
    return dict(model.encoder.state_dict())


def map_whisper_keys(encoder_keys, whisper_sd: Mapping[str, torch.Tensor]) -> Dict[str, str]:
    """encoder key -> whisper key for every tensor the reference copies (same key on both sides, positional embedding
    skipped: weight_init.py:57-66)."""
    return {k: k for k in encoder_keys if k != "positional_embedding" and k in whisper_sd}


def load_whisper_weights(encoder, whisper_model_name="openai/whisper-small", verbose=False, is_acoustic=False,
                         local_files_only=False):
    """Same signature and behaviour as the reference's `load_whisper_weights`; `encoder` is the `acoustic_encoder`
    sub-module of the B200 `AudioCodec` (a parameter holder: the kernels re-pack on the next forward).  Shapes must match
    (the reference's `copy_` would raise as well)."""
    whisper_sd = whisper_encoder_state_dict(whisper_model_name, local_files_only)
    state_dict = encoder.state_dict()
    mapping = map_whisper_keys(state_dict.keys(), whisper_sd)
    with torch.no_grad():
        for key, wkey in mapping.items():
            src = whisper_sd[wkey]
            if tuple(src.shape) != tuple(state_dict[key].shape):
                raise RuntimeError(f"load_whisper_weights: shape mismatch for {key}: {tuple(src.shape)} vs {tuple(state_dict[key].shape)}")
            state_dict[key].copy_(src.to(state_dict[key].dtype))
            if verbose:
                print(f"  ✓ {key}")
    encoder.load_state_dict(state_dict)
    owner = getattr(encoder, "_owner", None)
    if owner is not None:                          # invalidate the packed device copy of the owning AudioCodec
        owner._version += 1
    if verbose:
        print(f"Successfully loaded {len(mapping)} weight tensors")
        if is_acoustic:
            print("Note: Acoustic encoder loaded with Whisper weights (some modifications may apply)")
    return encoder

"""load_whisper_weights (reference utils/weight_init.py:11-76) — key mapping checked against what the reference's own
function overwrote when run on a random-init HF WhisperModel (tests/golden/whisper_init.json, made by
tests/golden/make_goldens_r2.py), on a synthetic HF-shaped encoder state dict.  No network, no GPU."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from simwhisper_codec_b200 import AudioCodec
from simwhisper_codec_b200.utils.weight_init import load_whisper_weights, map_whisper_keys, whisper_encoder_state_dict


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLDEN, "whisper_init.json")))


def _hf_like(golden, prefix=""):
    g = torch.Generator().manual_seed(5)
    return {prefix + k: torch.randn(shape, generator=g) for k, shape in golden["whisper_encoder_keys"].items()}


@pytest.mark.parametrize("prefix", ["", "encoder.", "model.encoder."])
def test_copies_exactly_the_keys_the_reference_copies(gen_params, sd_ex, golden, prefix):
    model = AudioCodec(gen_params)
    model.load_state_dict(sd_ex)
    src = _hf_like(golden, prefix)
    if prefix:
        src["decoder.embed_tokens.weight"] = torch.zeros(3, 3)          # decoder tensors of a full checkpoint are ignored
    before = {k: v.clone() for k, v in model.acoustic_encoder.state_dict().items()}
    v0 = model._version
    load_whisper_weights(model.acoustic_encoder, src, is_acoustic=True)
    after = model.acoustic_encoder.state_dict()
    changed = sorted(k for k in after if not torch.equal(after[k], before[k]))
    assert changed == golden["copied_keys"] and len(changed) == 186
    for k in changed:
        assert torch.equal(after[k], src[prefix + k])
    assert torch.equal(after["positional_embedding"], before["positional_embedding"])
    assert model._version > v0                                           # the packed device copy is rebuilt on the next call
    # the other sub-modules are untouched
    full = model.state_dict()
    assert all(torch.equal(full[k], sd_ex[k]) for k in full if not k.startswith("acoustic_encoder."))


def test_shape_mismatch_raises_like_copy_(gen_params, golden):
    model = AudioCodec(gen_params)
    src = _hf_like(golden)
    src["layers.3.fc1.weight"] = torch.zeros(7, 7)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        load_whisper_weights(model.acoustic_encoder, src)


def test_mapping_helpers(golden):
    sd = whisper_encoder_state_dict({"encoder.conv1.weight": torch.zeros(1), "decoder.x": torch.zeros(1), "proj_out.weight": torch.zeros(1)})
    assert list(sd) == ["conv1.weight"]
    m = map_whisper_keys(["positional_embedding", "conv1.weight", "extra"], {"positional_embedding": 0, "conv1.weight": 0})
    assert m == {"conv1.weight": "conv1.weight"}


def test_init_from_whisper_hook(gen_params, golden):
    gp = json.loads(json.dumps(gen_params))
    gp["acoustic_encoder"]["init_from_whisper"] = True
    model = AudioCodec(gp)
    model._init_whisper_weights()                                        # no path configured: warns and returns (model.py:66-68)
    model.whisper_model_path = _hf_like(golden)
    model._init_whisper_weights()
    assert float(model.acoustic_encoder.state_dict()["conv1.weight"].abs().sum()) > 0

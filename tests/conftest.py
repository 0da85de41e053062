import os
import sys

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def synthetic_wave(seed: int, n: int) -> torch.Tensor:
    """Same generator as tests/golden/make_goldens.py (0.1*randn clipped to +-1)."""
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(n, generator=g)).clamp_(-1, 1)


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def gen_params():
    cfg = yaml.safe_load(open(os.path.join(ROOT, "simwhisper_codec_b200", "config", "SimWhisperCodec.yaml")))
    return cfg["generator_params"]


@pytest.fixture(scope="session")
def sd_ex(gen_params):
    from simwhisper_codec_b200.weights import random_state_dict
    return random_state_dict(gen_params, seed=0, exercise=True)


@pytest.fixture(scope="session")
def sd_plain(gen_params):
    from simwhisper_codec_b200.weights import random_state_dict
    return random_state_dict(gen_params, seed=0, exercise=False)


def snr_db(ref: torch.Tensor, est: torch.Tensor) -> float:
    ref = ref.double().reshape(-1)
    est = est.double().reshape(-1)
    num = (ref ** 2).sum()
    den = ((ref - est) ** 2).sum()
    return float(10 * torch.log10(num / den.clamp_min(1e-300)))

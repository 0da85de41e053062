"""bench.py's CPU arm (`--impl reference`, `cpu_baseline`): it times the real reference when a copy of its tree is reachable
($SIMWHISPER_REF, baseline/_ref, /root/reference) and the pinned oracle port otherwise, on the same weights and inputs, and it
reports the median step with its spread.  No GPU."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("swc_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_synthetic_batch_is_seeded(bench):
    a, b = bench.synthetic_batch(2), bench.synthetic_batch(2)
    assert a.shape == (2, bench.WIN) and bool((a == b).all()) and float(a.abs().max()) <= 1.0
    assert not bool((bench.synthetic_batch(1, seed0=1001) == a[:1]).all())


def test_cpu_arm_runs_the_reference_when_its_tree_is_there(bench, monkeypatch):
    tree = bench.find_reference_tree()
    saved = list(sys.path), {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith(("utils.", "audiocodec"))}
    try:
        r = bench.cpu_reference_run(steps=1, warmup=0, items=1)
    finally:                      # the reference's top-level `audiocodec` / `utils` packages must not leak into other tests
        for k in [k for k in sys.modules if k == "utils" or k.startswith(("utils.", "audiocodec"))]:
            del sys.modules[k]
        sys.modules.update(saved[1])
        sys.path[:] = saved[0]
    assert r["kind"] == ("reference" if tree else "port")
    assert r["items"] == 1 and r["steps"] == 1 and r["value"] > 0.5 and r["seconds_per_step"] > 0
    assert "30 s windows per step" in r["sample"]
    # without a tree the port is timed
    monkeypatch.setattr(bench, "find_reference_tree", lambda: None)
    monkeypatch.setattr(bench, "synthetic_batch", lambda n, seed0=1000: bench.torch.zeros(n, 16000))
    monkeypatch.setattr(bench, "WIN", 16000)
    r2 = bench.cpu_reference_run(steps=2, warmup=0, items=1)
    assert r2["kind"] == "port" and r2["steps"] == 2 and r2["spread"] is not None

#!/usr/bin/env python
"""Where the wall time of encode()/decode() on BASELINE configs[3] goes: wall clock per call against the summed device time
of the library's kernels (swc_profile) - the difference is host work (planning, pageable uploads, launches, stitching)."""
import ctypes as C
import json
import os
import sys
import time

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simwhisper_codec_b200 import AudioCodec, _lib  # noqa: E402
from simwhisper_codec_b200.weights import random_state_dict  # noqa: E402


def main():
    gp = yaml.safe_load(open(os.path.join(ROOT, "simwhisper_codec_b200", "config", "SimWhisperCodec.yaml")))["generator_params"]
    precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    model = AudioCodec(gp, precision=precision, max_batch=256)
    model.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
    lib = _lib.load()
    gl = torch.Generator().manual_seed(123)
    lens = [int(16000 * (2 + 28 * float(torch.rand((), generator=gl)))) for _ in range(256)]
    g = torch.Generator().manual_seed(7)
    wavs = [(0.1 * torch.randn(n, generator=g)).clamp_(-1, 1) for n in lens]
    pinned = [w.pin_memory() for w in wavs]
    ms_cls, n_cls = (C.c_double * 8)(), (C.c_int64 * 8)()
    for name, src in (("pageable", wavs), ("pinned", pinned)):
        for rep in range(3):
            torch.cuda.synchronize()
            lib.swc_profile(1)
            t0 = time.perf_counter()
            codes = model.encode(src)["codes_list"]
            t_launch = time.perf_counter() - t0
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            lib.swc_profile_read(ms_cls, n_cls, 8)
            enc_dev = sum(ms_cls)
            lib.swc_profile(1)
            out = model.decode(codes)["syn_wav_list"]
            t_launch_d = time.perf_counter() - t1
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            lib.swc_profile_read(ms_cls, n_cls, 8)
            dec_dev = sum(ms_cls)
            lib.swc_profile(0)
        print(json.dumps({"input": name, "precision": precision, "audio_s": round(sum(lens) / 16000, 1),
                          "encode_wall_ms": round((t1 - t0) * 1e3, 1), "encode_host_until_return_ms": round(t_launch * 1e3, 1),
                          "encode_kernels_ms": round(enc_dev, 1), "decode_wall_ms": round((t2 - t1) * 1e3, 1),
                          "decode_host_until_return_ms": round(t_launch_d * 1e3, 1), "decode_kernels_ms": round(dec_dev, 1),
                          "audio_s_per_s": round(sum(lens) / 16000 / (t2 - t0), 1)}), flush=True)


if __name__ == "__main__":
    main()

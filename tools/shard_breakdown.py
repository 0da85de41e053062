#!/usr/bin/env python
"""What ONE rank of an N-GPU `ShardedCodec` job does, timed on a single GPU (no collectives): the share of BASELINE
configs[3] that `windows.shard_round_robin` deals to rank 0 of `world` ranks goes through the same
`encode_jobs` / `decode_jobs` calls, with host time stamps at every phase boundary and the summed device time of the
library's kernels (swc_profile) next to the wall clock.  The difference is what strong scaling loses to the host.

usage: shard_breakdown.py [world=8] [precision=bf16] [api_chunk]"""
import ctypes as C
import json
import os
import sys
import time

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simwhisper_codec_b200 import AudioCodec, _lib, windows  # noqa: E402
from simwhisper_codec_b200.weights import random_state_dict  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    gp = yaml.safe_load(open(os.path.join(ROOT, "simwhisper_codec_b200", "config", "SimWhisperCodec.yaml")))["generator_params"]
    model = AudioCodec(gp, precision=precision, max_batch=256)
    if len(sys.argv) > 3:
        model.api_chunk = int(sys.argv[3])
    model.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
    lib = _lib.load()
    dev = torch.device("cuda")
    gl = torch.Generator().manual_seed(123)
    lens = [int(16000 * (2 + 28 * float(torch.rand((), generator=gl)))) for _ in range(256)]
    g = torch.Generator().manual_seed(7)
    wavs = [(0.1 * torch.randn(n, generator=g)).clamp_(-1, 1) for n in lens]
    if os.environ.get("PINNED", "0") == "1":
        wavs = [w.pin_memory() for w in wavs]
    ms_cls, n_cls = (C.c_double * 8)(), (C.c_int64 * 8)()
    m = model
    prof = os.environ.get("NOPROF", "0") == "0"
    for rep in range(4):
        torch.cuda.synchronize()
        lib.swc_profile(1 if prof else 0)
        st = {}
        t0 = time.perf_counter()
        jobs = windows.plan_encode(lens, 10, m.input_sample_rate, m.max_audio_seconds)
        shards = windows.shard_round_robin(len(jobs), [j.n_valid for j in jobs], world)
        st["enc_plan"] = time.perf_counter() - t0
        mine = m.encode_jobs(wavs, [jobs[j] for j in shards[int(os.environ.get('RANK_EMU', '0'))]], dev)
        st["enc_jobs_host"] = time.perf_counter() - t0
        torch.cuda.synchronize()
        st["enc_synced"] = time.perf_counter() - t0
        lib.swc_profile_read(ms_cls, n_cls, 8)
        st["enc_kernels"] = sum(ms_cls) / 1e3
        # the other ranks' codes: this rank's own rows stand in for them (same shapes, same stitching work)
        full = mine[:, torch.arange(len(jobs), device=dev) % mine.shape[1]]
        torch.cuda.synchronize()
        lib.swc_profile(1 if prof else 0)
        t1 = time.perf_counter()
        codes_list = m.stitch_codes(full.contiguous(), lens, jobs, 10)
        st["stitch_codes_host"] = time.perf_counter() - t1
        clens = [int(c.shape[-1]) for c in codes_list]
        groups = windows.plan_decode(clens, 10, m.input_sample_rate, m.max_audio_seconds, m.encoder_downsample_rate)
        st["dec_plan"] = time.perf_counter() - t1
        n_my = 0
        order = sorted(groups.items())
        all_sh = windows.shard_groups([[j.n_valid for j in dj] for _, dj in order], world)
        which = int(os.environ.get("RANK_EMU", "0"))
        for (pad_len, dj), sh in zip(order, all_sh):
            my = [dj[j] for j in sh[which]]
            n_my += len(my)
            if my:
                m.decode_jobs(codes_list, my, dev)
        st["dec_jobs_host"] = time.perf_counter() - t1
        torch.cuda.synchronize()
        st["dec_synced"] = time.perf_counter() - t1
        lib.swc_profile_read(ms_cls, n_cls, 8)
        st["dec_kernels"] = sum(ms_cls) / 1e3
        lib.swc_profile(0)
        if rep:
            print(json.dumps({"world": world, "precision": precision, "api_chunk": m.api_chunk, "pinned": os.environ.get("PINNED", "0"), "enc_windows": len(shards[0]),
                              "dec_windows": n_my, **{k: round(v * 1e3, 2) for k, v in st.items()}}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""encode()/decode() API on variable-length and long-form batches, window-sharded over the GPUs of one node
(BASELINE.json configs[3] and [4]).  Launch with torchrun (one rank per GPU) or plain python (1 GPU).

  config 4: N items with lengths U(2, 30) s (seeded), padded + masked, sharded by windows.ShardedCodec
  config 5: M long-form items (default 10 min) chunked into 30 s windows every 20 s, keep-first stitching

Checks on rank 0 (on a subset): sharded codes bit-equal to the single-GPU API, sharded waveforms equal to the single-GPU
API's (same global T').  Prints one JSON line per config: valid audio-seconds per second (encode + decode, whole job).
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simwhisper_codec_b200 import AudioCodec  # noqa: E402
from simwhisper_codec_b200.parallel import ShardedCodec  # noqa: E402
from simwhisper_codec_b200.weights import random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=256)
    ap.add_argument("--long-items", type=int, default=8)
    ap.add_argument("--long-seconds", type=int, default=600)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--check", type=int, default=6, help="items compared against the single-GPU API")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--pinned", type=int, default=1, help="page-lock the input tensors (as bench.py's sharded_api leg does)")
    ap.add_argument("--sync-between", type=int, default=0, help="synchronise between encode() and decode() (splits the time)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    gp = yaml.safe_load(open(os.path.join(ROOT, "simwhisper_codec_b200", "config", "SimWhisperCodec.yaml")))["generator_params"]
    model = AudioCodec(gp, precision=args.precision, max_batch=128)
    model.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
    sc = ShardedCodec(model)

    def run(name, lens):
        g = torch.Generator().manual_seed(7)
        wavs = [(0.1 * torch.randn(n, generator=g)).clamp_(-1, 1) for n in lens]
        # correctness on a subset (every rank takes part in the sharded calls)
        if args.pinned:
            wavs = [w.pin_memory() for w in wavs]
        sub = wavs[: args.check]
        c_sh = sc.encode(sub, device=dev)["codes_list"]
        w_sh = sc.decode(c_sh, device=dev)["syn_wav_list"]          # stitched on rank 0 only
        ok = None
        if rank == 0:
            c_1 = model.encode(sub, device=dev)["codes_list"]
            w_1 = model.decode(c_1, device=dev)["syn_wav_list"]
            ok = all(torch.equal(a, b) for a, b in zip(c_sh, c_1)) and all(torch.equal(a, b) for a, b in zip(w_sh, w_1))
        secs = sum(lens) / 16000.0
        best, all_dt = None, []
        for _ in range(args.reps + 1):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            codes = sc.encode(wavs, device=dev)["codes_list"]
            if args.sync_between:
                torch.cuda.synchronize()
            t1 = time.perf_counter()
            out = sc.decode(codes, device=dev)["syn_wav_list"]
            torch.cuda.synchronize()
            dist.barrier()
            dt = time.perf_counter() - t0
            all_dt.append(round(dt * 1e3, 2))
            if best is None or dt < best:
                best, split = dt, (t1 - t0, time.perf_counter() - t1)
        if rank == 0:
            n_codes = sum(int(c.shape[-1]) for c in codes)
            print(json.dumps({"config": name, "n_gpus": world, "items": len(lens), "audio_seconds": round(secs, 1),
                              "windows_encode": sum((n + 319999) // 320000 for n in lens), "code_frames": n_codes,
                              "wall_s": round(best, 4), "all_ms": all_dt, "joint_sharding": os.environ.get("SWC_SHARD_JOINT", "1"), "encode_s": round(split[0], 3), "decode_s": round(split[1], 3), "audio_s_per_s": round(secs / best, 1), "precision": args.precision,
                              "sharded_equals_single_gpu": ok, "timing": "wall clock incl. host window planning, H2D of "
                              "the utterances, NCCL gather of codes and waveforms; best of %d" % args.reps}), flush=True)

    g = torch.Generator().manual_seed(123)
    lens4 = [int(16000 * (2 + 28 * float(torch.rand((), generator=g)))) for _ in range(args.items)]
    run("configs[3]: variable-length 2-30 s", lens4)
    if args.long_items > 0:
        run(f"configs[4]: long-form {args.long_seconds} s", [16000 * args.long_seconds] * args.long_items)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the SimWhisper-Codec hot path on B200 (contract: see the task's bench.py section).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    the reference algorithm on the host cores
                                                           (oracle/port.py — the reference itself is
                                                           pure Python and is not on the GPU box)

Metric (BASELINE.json): audio-seconds per second of full-codec encode->quantize->decode at 16 kHz.
Workload at any N: BASELINE.json configs[2] — per GPU a batch of 256 x 30 s synthetic utterances,
random-init weights of config/SimWhisperCodec.yaml, bf16 tensor-core mode, one single-pass
inference_tokenize -> inference_detokenize per step ("weak" scaling: 256 windows per GPU).
`value` times the step with the batch already in HBM; `e2e` times the public API call with HOST
(pinned) buffers, host->device and device->host copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import yaml  # noqa: E402

SR = 16000
WIN = 480000
# algorithmic work per 30 s window (SURVEY.md 8d / BASELINE.md 3), GFLOP, 2*MAC, nothing skipped
GF_TOTAL = 1311.86
GF_ATTN = 24 * 6.912
GF_MEL = 1.061
GF_IDFT = 2.465
# DRAM traffic of the tcgen05 GEMM class per window, from the committed ncu launch list (profiles/r2_ncu_launches.md)
GEMM_DRAM_GB_PER_WINDOW = 2.65
GF_TC_GEMM = GF_TOTAL - GF_ATTN - 0.096              # contractions that run on the tcgen05 GEMM kernels: everything but
                                                       # attention and the 80-bin mel filterbank (the forward and inverse DFTs
                                                       # run as split-bf16 GEMMs; their algorithmic FLOPs are counted once)


def gen_params():
    return yaml.safe_load(open(os.path.join(ROOT, "simwhisper_codec_b200", "config", "SimWhisperCodec.yaml")))["generator_params"]


def synthetic_batch(n, seed0=1000):
    g = torch.Generator().manual_seed(seed0)
    return (0.1 * torch.randn(n, WIN, generator=g)).clamp_(-1, 1)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, pw, mx, reasons = [], [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                pw.append(float(r[3]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # samples under load = those drawing at least 70 % of the highest power seen (idle samples read the boost clock)
        top = max(pw) if pw else 0.0
        busy = sorted(c for c, w in zip(sm, pw) if w >= 0.7 * top)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(busy),
                "power_w_max": top or None}


def find_reference_tree():
    """The real reference, if a copy is reachable on this box: $SIMWHISPER_REF, baseline/_ref, /root/reference."""
    for cand in (os.environ.get("SIMWHISPER_REF"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "audiocodec", "model.py")):
            return cand
    return None


def cpu_reference_run(steps, warmup, items=4):
    """The reference algorithm in fp32 on all host threads.  One step = the same call sequence as the GPU arm's step
    (single-pass inference_tokenize -> inference_detokenize) on `items` x 30 s windows (B = 4; 256
    windows would take half an hour of CPU time per step).  The real reference is used when a copy of its tree is reachable
    (kind "reference"), else the pinned oracle port (kind "port").  Returns the MEDIAN step."""
    from simwhisper_codec_b200.weights import random_state_dict
    torch.set_num_threads(os.cpu_count())
    gp = gen_params()
    sd = random_state_dict(gp, seed=0, exercise=True)
    x = synthetic_batch(items)[:, None, :]
    lens = torch.full((items,), WIN, dtype=torch.long)
    ref_tree = find_reference_tree()
    kind = "port"
    if ref_tree is not None:
        try:
            import importlib
            import warnings
            warnings.filterwarnings("ignore")
            sys.dont_write_bytecode = True
            sys.path.insert(0, ref_tree)
            for name in [m for m in sys.modules if m == "audiocodec" or m.startswith("audiocodec.") or m == "utils" or m.startswith("utils.")]:
                del sys.modules[name]
            RefCodec = importlib.import_module("audiocodec.model").AudioCodec
            ref = RefCodec(yaml.safe_load(open(os.path.join(ref_tree, "config", "SimWhisperCodec.yaml")))["generator_params"]).eval()
            ref.load_state_dict(sd, strict=True)
            kind = "reference"
        except Exception as e:      # an unusable copy: fall back to the port, say so
            print(f"[bench] reference tree at {ref_tree} not usable ({str(e)[:120]}); timing the oracle port", file=sys.stderr)
            sys.path.remove(ref_tree)
    if kind == "reference":
        def step():
            with torch.inference_mode():
                r = ref.inference_tokenize(x, lens)
                ref.inference_detokenize(r["codes"], r["codes_lengths"])
    else:
        from oracle import port

        def step():
            with torch.inference_mode():
                r = port.tokenize(sd, x, lens)
                port.detokenize(sd, r["codes"], r["codes_lengths"])

    for _ in range(warmup):
        step()
    times = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": items * 30.0 / med, "seconds_per_step": med, "kind": kind, "items": items, "steps": len(times),
            "spread": (times[-1] - times[0]) / med if len(times) > 1 else None,
            "sample": f"{items} x 30 s windows per step (single-pass inference_tokenize -> inference_detokenize, fp32), median of "
                      f"{len(times)} step(s) after {warmup} warm-up, "
                      + ("the reference's own AudioCodec on CPU" if kind == "reference" else "oracle/port.py (the reference's ATen ops restated)")
                      + f", {os.cpu_count()} host threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="30 s windows per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3"])
    ap.add_argument("--max-batch", type=int, default=int(os.environ.get("SWC_MAX_BATCH", 256)),
                    help="windows per kernel launch (the 256-window step runs as 256/max_batch sub-batches)")
    ap.add_argument("--e2e-chunk", type=int, default=256,
                    help="windows per chunk of the end-to-end step (a chunk's copies overlap the compute of the neighbouring "
                         "chunks and steps; one chunk per step measured best: 400 vs 413 ms with two)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-leg", action="store_true",
                    help="skip the extra measurement of the parity-grade precision (bf16x3) at N = 1")
    ap.add_argument("--no-check", action="store_true", help="skip the output check of the timed configuration against bf16x3")
    ap.add_argument("--sharded-api", default="on", choices=["on", "off"],
                    help="also run ShardedCodec.encode+decode on BASELINE configs[3] (fixed job: strong scaling over N) and a "
                         "configs[4] scaled to 4 ten-minute items per GPU")
    args = ap.parse_args()

    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    config = {"workload": f"BASELINE.json configs[2]: full codec encode->decode, batch {args.batch} x 30 s per GPU, "
                          f"{args.precision}, single-pass inference_tokenize->inference_detokenize, "
                          "random-init SimWhisperCodec.yaml weights",
              "windows_per_gpu": args.batch, "window_seconds": 30, "sample_rate": SR,
              "l2_policy": "inputs (492 MB) and activations (GBs) exceed the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        # a bounded sample of the same workload: 4 of the 256 windows per step, fp32, all host threads; the run is sized to end
        # within a few minutes whatever --steps says (a step is ~5 s of CPU time on 16 cores)
        items = 4
        steps = max(3, min(args.steps, 12))
        warm = max(1, min(args.warmup, 2))
        r = cpu_reference_run(steps, warm, items)
        ref_config = dict(config)
        ref_config["workload"] = (f"BASELINE.json configs[2] on the host CPU: the same single-pass inference_tokenize->inference_detokenize, "
                                  f"{items} x 30 s windows per step (a bounded sample of the 256-window batch), fp32, random-init "
                                  "SimWhisperCodec.yaml weights")
        ref_config["windows_per_step"] = items
        ref_config["precision"] = "fp32"
        del ref_config["windows_per_gpu"]
        emit({
            "impl": "reference", "metric": "audio-sec/sec encode+decode (16 kHz)", "value": r["value"], "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": r["steps"], "warmup": warm, "ms_per_step": r["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": ref_config,
            "cpu_baseline": {"value": r["value"], "unit": "audio-s/s", "cores": os.cpu_count(), "kind": r["kind"],
                             "sample": r["sample"], "spread": r["spread"]},
            "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return

    import torch.distributed as dist
    from simwhisper_codec_b200 import AudioCodec, _lib
    from simwhisper_codec_b200.weights import random_state_dict
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # (NCCL's banner / NCCL_DEBUG output cannot reach the JSON line: fd 1 points at stderr until emit())
        dist.init_process_group("nccl", device_id=dev)

    gp = gen_params()
    model = AudioCodec(gp, precision=args.precision, max_batch=args.max_batch)
    model.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
    B = args.batch
    host_x = synthetic_batch(B, seed0=1000 + rank).pin_memory()
    host_y = torch.empty(B, 1280 * 375, dtype=torch.float32).pin_memory()
    host_codes = torch.empty(8, B, 375, dtype=torch.int32).pin_memory()
    lens = torch.full((B,), WIN, dtype=torch.int64, device=dev)
    x_dev = host_x.to(dev)[:, None, :]
    lib = _lib.load()

    def step_resident():
        r = model.inference_tokenize(x_dev, lens)
        return model.inference_detokenize(r["codes"], r["codes_lengths"])

    # end to end through the public API with HOST buffers: the batch goes through in `e2e_chunk`-window chunks, each
    # chunk's host->device copy, compute and device->host copy on their own streams so copies overlap the other chunk's compute
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    e2e_chunk = max(1, min(args.e2e_chunk, args.max_batch))
    chunks = [slice(s0, min(s0 + e2e_chunk, B)) for s0 in range(0, B, e2e_chunk)]

    def step_e2e():
        # Only event dependencies tie the three streams together, so consecutive steps pipeline as well: the next step's
        # host->device copy runs under this step's compute, this step's device->host copy under the next step's compute
        # (the caching allocator orders buffer reuse across streams through record_stream).
        main = torch.cuda.current_stream()
        staged = []
        for sl in chunks:
            with torch.cuda.stream(copy_in):
                xd = host_x[sl].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_in)
            staged.append((xd, ev))
        for sl, (xd, ev) in zip(chunks, staged):
            main.wait_event(ev)
            xd.record_stream(main)
            r = model.inference_tokenize(xd[:, None, :], lens[sl])
            out = model.inference_detokenize(r["codes"], r["codes_lengths"])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                host_codes[:, sl].copy_(r["codes"], non_blocking=True)
                host_y[sl].copy_(out["y"][:, 0], non_blocking=True)
            r["codes"].record_stream(copy_out)
            out["y"].record_stream(copy_out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        main = torch.cuda.current_stream()
        main.wait_stream(copy_out)          # the timed region ends when the last device->host copy has landed
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    lib.swc_profile(0)
    ms = timed(step_resident, args.steps)
    ms_cls = (C.c_double * 8)()
    n_cls = (C.c_int64 * 8)()
    lib.swc_profile_read(ms_cls, n_cls, 8)
    launches = int(sum(n_cls))
    clocks = sampler.stop()
    value = world * B * 30.0 / (ms * 1e-3)

    # per-kernel-class device time over the same step with CUDA events around every launch.  An un-profiled step runs
    # directly in front of it (no synchronisation in between), so the profiled step sees the steady-state clocks of the
    # power-capped timed region rather than the boost of a GPU that has just been idle.
    barrier()
    step_resident()
    lib.swc_profile(1)
    step_resident()
    torch.cuda.synchronize()
    lib.swc_profile_read(ms_cls, n_cls, 8)
    lib.swc_profile(0)
    cls_ms = {k: float(ms_cls[i]) for i, k in enumerate(_lib.KCLASS)}
    cls_n = {k: int(n_cls[i]) for i, k in enumerate(_lib.KCLASS)}
    tot_cls = sum(cls_ms.values()) or 1.0

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = world * B * 30.0 / (ms_e2e * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PF/s sustained (B200_PROFILING.md)"
    gemm_ms = cls_ms["gemm_tcgen05"]
    achieved = (GF_TC_GEMM * B / 1e3) / (gemm_ms * 1e-3) if gemm_ms > 0 else 0.0
    roofline = {"kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05 bf16 GEMMs, all instantiations)", "bound": "tensor",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                "traffic": GEMM_DRAM_GB_PER_WINDOW * 1e9 * B / max(cls_n["gemm_tcgen05"], 1),
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of the gemm_tc* launches of one "
                                "step under ncu, profiles/r2_ncu_launches.md, averaged per launch)",
                "peak_source": peak_src,
                "algorithmic_gflop_per_window": GF_TC_GEMM, "windows_per_step": B, "launches_per_step": cls_n["gemm_tcgen05"],
                "avg_launch_ms": gemm_ms / max(cls_n["gemm_tcgen05"], 1),
                "class_ms_per_step": cls_ms, "class_share": {k: v / tot_cls for k, v in cls_ms.items()},
                "class_launches": cls_n}
    def snr_db(ref, est):
        ref, est = ref.double().reshape(-1), est.double().reshape(-1)
        return float(10 * torch.log10((ref ** 2).sum() / ((ref - est) ** 2).sum().clamp_min(1e-300)))

    # ---- the timed configuration's own output, checked on this device against the parity-grade mode (bf16x3, which the
    #      GPU tests hold to the fp32 bars against the reference): 8 of the 256 windows, codes and waveform
    check = None
    if not args.no_check and args.precision == "bf16":
        pick = sorted({int(round(i * (B - 1) / 7)) for i in range(8)})
        r_all = model.inference_tokenize(x_dev, lens)                       # the timed step's call, outputs kept this time
        y_all = model.inference_detokenize(r_all["codes"], r_all["codes_lengths"])["y"]
        codes16 = r_all["codes"][:, pick].clone()
        finite = bool(torch.isfinite(y_all).all())
        rms = float(y_all[pick].float().pow(2).mean().sqrt())
        del r_all, y_all
        mx = AudioCodec(gp, precision="bf16x3", max_batch=len(pick))
        mx.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
        xs = host_x[pick].to(dev)[:, None, :]
        r3 = mx.inference_tokenize(xs, lens[pick])
        flip = float((codes16 != r3["codes"]).float().mean())
        y3 = mx.inference_detokenize(r3["codes"], r3["codes_lengths"])["y"]
        y16 = model.inference_detokenize(r3["codes"], r3["codes_lengths"])["y"]   # identical codes through the timed mode
        snr = snr_db(y3, y16)
        check = {"windows": pick, "flip_rate": flip, "snr_db": snr, "finite": finite, "rms": rms,
                 "bounds": {"flip_rate_max": 0.05, "snr_db_min": 38.0},
                 "ok": bool(finite and flip <= 0.05 and snr >= 38.0),
                 "what": "codes of the timed bf16 step vs bf16x3 on the same inputs; bf16 vs bf16x3 waveform on identical codes"}
        del mx, xs, r3, y3, y16
        torch.cuda.empty_cache()

    # ---- BASELINE configs[3] / [4] through the sharded public API (windows dealt to the ranks, codes all-gathered, waveforms
    #      gathered to rank 0): valid audio-seconds per second, host planning and copies included
    sharded_api = None
    if args.sharded_api == "on" and args.precision == "bf16":
        from simwhisper_codec_b200.parallel import ShardedCodec
        if world == 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29577")
            dist.init_process_group("gloo", rank=0, world_size=1)
        sc = ShardedCodec(model)

        def run_api(lens_s, reps=2):
            g = torch.Generator().manual_seed(7)
            wavs = [(0.1 * torch.randn(n, generator=g)).clamp_(-1, 1) for n in lens_s]
            sub = wavs[:8]                                   # correctness on a subset: every rank takes part
            c_sh = sc.encode(sub, device=dev)["codes_list"]
            w_sh = sc.decode(c_sh, device=dev)["syn_wav_list"]
            ok = None
            if rank == 0:
                c_1 = model.encode(sub, device=dev)["codes_list"]
                w_1 = model.decode(c_1, device=dev)["syn_wav_list"]
                ok = all(torch.equal(a, b) for a, b in zip(c_sh, c_1)) and all(torch.equal(a, b) for a, b in zip(w_sh, w_1))
            def timed_api(inputs, n):
                best = None
                for _ in range(n):
                    barrier()
                    t0 = time.perf_counter()
                    codes = sc.encode(inputs, device=dev)["codes_list"]
                    sc.decode(codes, device=dev)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    t = torch.tensor([dt], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    best = float(t[0]) if best is None else min(best, float(t[0]))
                return best

            pinned = [w.pin_memory() for w in wavs]          # as the e2e leg: the caller's buffers are page-locked
            best = timed_api(pinned, reps + 1)               # the first pass warms the workspaces
            pageable = timed_api(wavs, 2)
            secs = sum(lens_s) / SR
            return {"value": secs / best, "unit": "valid audio-s/s", "items": len(lens_s), "audio_seconds": secs,
                    "windows": sum((n + 319999) // 320000 for n in lens_s), "wall_s": best, "equals_single_gpu": ok,
                    "value_pageable_inputs": secs / pageable}

        gl = torch.Generator().manual_seed(123)
        lens3 = [int(SR * (2 + 28 * float(torch.rand((), generator=gl)))) for _ in range(256)]
        long_items = 4 * world
        sharded_api = {"configs[3]": run_api(lens3), "configs[4]": run_api([SR * 600] * long_items),
                       "note": "ShardedCodec.encode + decode (default overlap 10 s), wall clock incl. host planning, H2D from pinned host "
                               "tensors (value_pageable_inputs: the same from ordinary pageable tensors), NCCL gather of codes (all "
                               "ranks) and of the kept samples (rank 0); configs[3] = 256 items of 2-30 s (strong scaling: fixed "
                               f"job), configs[4] scaled to {long_items} items of 10 min (4 per GPU, 30 windows each); best of 2"}

    # ---- the same single pass in the parity-grade precision (fp32 activations, three-product split-bf16 contractions and
    #      attention on tcgen05: index flips < 0.1 %, decode SNR > 90 dB vs the reference) at the FULL batch, with its own
    #      class timings and roofline: N = 1 only
    parity_mode = None
    if world == 1 and not args.no_parity_leg and args.precision == "bf16":
        try:
            del x_dev
            model._native = None
            torch.cuda.empty_cache()
            mx = AudioCodec(gp, precision="bf16x3", max_batch=args.max_batch)
            mx.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
            xx = host_x.to(dev)[:, None, :]

            def step_x3():
                r = mx.inference_tokenize(xx, lens)
                return mx.inference_detokenize(r["codes"], r["codes_lengths"])

            for _ in range(2):
                step_x3()
            sampler3 = ClockSampler(local)
            sampler3.start()
            time.sleep(0.2)
            ms_x3 = timed(step_x3, max(2, min(args.steps, 3)))
            clocks3 = sampler3.stop()
            step_x3()
            lib.swc_profile(1)
            step_x3()
            torch.cuda.synchronize()
            lib.swc_profile_read(ms_cls, n_cls, 8)
            lib.swc_profile(0)
            cls3 = {k: float(ms_cls[i]) for i, k in enumerate(_lib.KCLASS)}
            g3 = cls3["gemm_tcgen05"]
            # three bf16 products per contraction: the tensor pipe does 3x the algorithmic FLOPs
            alg = (GF_TC_GEMM * B / 1e3) / (g3 * 1e-3) if g3 > 0 else 0.0
            parity_mode = {"precision": "bf16x3", "value": B * 30.0 / (ms_x3 * 1e-3), "unit": "audio-s/s", "windows": B,
                           "ms_per_step": ms_x3, "clocks": clocks3, "class_ms_per_step": cls3,
                           "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel, three-product (hi|lo) operands",
                                        "achieved_algorithmic": alg, "achieved": 3 * alg, "peak": peak_tf, "unit": "TFLOP/s",
                                        "frac": 3 * alg / peak_tf if peak_tf else None,
                                        "note": "achieved = bf16 tensor-pipe work (3 products per contraction) / GEMM-class time; "
                                                "achieved_algorithmic counts every contraction once"},
                           "attention_tflops_algorithmic": (GF_ATTN * B / 1e3) / (cls3["attention"] * 1e-3) if cls3["attention"] > 0 else None,
                           "note": "fp32 activations, dense contractions and attention as three bf16 products with fp32 accumulation on "
                                   "tcgen05; meets the fp32 parity bars (tests/test_gpu_parity.py::test_bf16x3_mode_meets_the_fp32_bars, "
                                   "test_config1_all_64_windows)"}
            del mx, xx
            torch.cuda.empty_cache()
        except Exception as e:      # informational leg: never fail the bench line
            parity_mode = {"precision": "bf16x3", "error": str(e)[:200]}
    if rank != 0:
        if dist.is_initialized():
            dist.destroy_process_group()
        return
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N = 1 only
        r = cpu_reference_run(3, 1, 4)
        cpu_baseline = {"value": r["value"], "unit": "audio-s/s", "cores": os.cpu_count(), "kind": r["kind"],
                        "sample": r["sample"], "spread": r["spread"]}
    out = {
        "metric": "audio-sec/sec encode+decode (16 kHz)", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config,
        "x_realtime_per_gpu": value / world,
        "e2e": {"value": e2e, "unit": "audio-s/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": host_x.numel() * 4, "d2h_bytes_per_step": host_y.numel() * 4 + host_codes.numel() * 4,
                "api": f"AudioCodec.inference_tokenize -> inference_detokenize per {e2e_chunk}-window chunk from pinned host buffers; "
                       "H2D / compute / D2H of consecutive chunks and steps overlap on three streams"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "check": check, "parity_mode": parity_mode, "sharded_api": sharded_api,
    }
    emit(out)
    if dist.is_initialized():
        dist.destroy_process_group()
    if check is not None and not check["ok"]:
        raise SystemExit(f"bench.py: the timed configuration's output failed its check: {check}")


if __name__ == "__main__":
    main()

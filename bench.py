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
# DRAM traffic of the tcgen05 GEMM class per window, from the committed ncu launch list (profiles/r1_ncu_launches_v7_final.md)
GEMM_DRAM_GB_PER_WINDOW = 2.66
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


def cpu_reference_run(steps, warmup, windows_per_step=1):
    """Reference algorithm (oracle port, fp32) on all host threads; one step = `windows_per_step` windows."""
    from oracle import port
    from simwhisper_codec_b200.weights import random_state_dict
    torch.set_num_threads(os.cpu_count())
    sd = random_state_dict(gen_params(), seed=0, exercise=True)
    x = synthetic_batch(windows_per_step)[:, None, :]
    lens = torch.full((windows_per_step,), WIN, dtype=torch.long)

    def step():
        with torch.inference_mode():
            r = port.tokenize(sd, x, lens)
            port.detokenize(sd, r["codes"], r["codes_lengths"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return windows_per_step * 30.0 / dt, dt


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
        warm = min(args.warmup, 1)
        v, dt = cpu_reference_run(args.steps, warm)
        emit({
            "impl": "reference", "metric": "audio-sec/sec encode+decode (16 kHz)", "value": v, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "1 x 30 s window per step (tokenize+detokenize), oracle/port.py on all host threads"},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
        # rank 0 must print exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION/INFO) off stdout
        if not os.environ.get("SWC_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
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
                                "step under ncu, profiles/r1_ncu_launches_v7_final.md, averaged per launch)",
                "peak_source": peak_src,
                "algorithmic_gflop_per_window": GF_TC_GEMM, "windows_per_step": B, "launches_per_step": cls_n["gemm_tcgen05"],
                "avg_launch_ms": gemm_ms / max(cls_n["gemm_tcgen05"], 1),
                "class_ms_per_step": cls_ms, "class_share": {k: v / tot_cls for k, v in cls_ms.items()},
                "class_launches": cls_n}
    # the same single pass in the parity-grade precision (fp32 activations, three-product split-bf16 contractions: index flips
    # < 0.1 %, decode SNR > 90 dB vs the reference): informational, N = 1 only, 64 windows resident in HBM
    parity_mode = None
    if world == 1 and not args.no_parity_leg and args.precision == "bf16":
        try:
            del x_dev
            model._native = None
            torch.cuda.empty_cache()
            pb = min(64, B)
            mx = AudioCodec(gp, precision="bf16x3", max_batch=pb)
            mx.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
            xx = host_x[:pb].to(dev)[:, None, :]
            ll = lens[:pb]

            def step_x3():
                r = mx.inference_tokenize(xx, ll)
                return mx.inference_detokenize(r["codes"], r["codes_lengths"])

            for _ in range(2):
                step_x3()
            ms_x3 = timed(step_x3, 2)
            parity_mode = {"precision": "bf16x3", "value": pb * 30.0 / (ms_x3 * 1e-3), "unit": "audio-s/s", "windows": pb,
                           "ms_per_step": ms_x3,
                           "note": "fp32 activations, dense contractions and attention as three bf16 products with fp32 "
                                   "accumulation; meets the fp32 parity bars (tests/test_gpu_parity.py::test_bf16x3_mode_meets_the_fp32_bars)"}
            del mx, xx
            torch.cuda.empty_cache()
        except Exception as e:      # informational leg: never fail the bench line
            parity_mode = {"precision": "bf16x3", "error": str(e)[:200]}
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N = 1 only
        v, dt = cpu_reference_run(1, 1)
        cpu_baseline = {"value": v, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": "1 x 30 s window (tokenize+detokenize, fp32) after 1 warm-up, oracle/port.py, all host threads"}
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
        "parity_mode": parity_mode,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

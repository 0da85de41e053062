#!/usr/bin/env python
"""Latency of one encode->decode call at small batch (serving shapes), eager launches against a captured CUDA graph.

At 256 windows the ~880 kernel launches of a step are hidden behind 400 ms of device time; at one or a few windows the
step is launch-bound on the host.  The library makes no allocation, no synchronisation and no host read-back inside a call
and the wrapper keeps one persistent workspace, so a whole inference_tokenize -> inference_detokenize call captures into a
CUDA graph as it is (`torch.cuda.graph`); replaying it removes the launch overhead.  Prints one JSON line per shape."""
import json
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simwhisper_codec_b200 import AudioCodec  # noqa: E402
from simwhisper_codec_b200.weights import random_state_dict  # noqa: E402


def main():
    gp = yaml.safe_load(open(os.path.join(ROOT, "simwhisper_codec_b200", "config", "SimWhisperCodec.yaml")))["generator_params"]
    precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    model = AudioCodec(gp, precision=precision)
    model.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
    reps = 20
    for B, secs in [(1, 10), (1, 30), (4, 30), (16, 30), (64, 30)]:
        g = torch.Generator().manual_seed(B * 100 + secs)
        x = (0.1 * torch.randn(B, 1, secs * 16000, generator=g)).cuda()
        lens = torch.full((B,), secs * 16000, dtype=torch.int64, device="cuda")

        def call():
            r = model.inference_tokenize(x, lens)
            return r["codes"], model.inference_detokenize(r["codes"], r["codes_lengths"])["y"]

        for _ in range(3):
            codes0, y0 = call()
        torch.cuda.synchronize()

        def timed(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_eager = timed(call)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            call()                                         # warm-up on the capture stream
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(graph):
            codes_g, y_g = call()
        graph.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(codes_g, codes0) and torch.equal(y_g, y0))
        ms_graph = timed(graph.replay)
        print(json.dumps({"batch": B, "seconds": secs, "precision": precision, "ms_eager": round(ms_eager, 3),
                          "ms_graph": round(ms_graph, 3), "x_realtime_graph": round(B * secs / (ms_graph * 1e-3), 1),
                          "graph_equals_eager": same}), flush=True)


if __name__ == "__main__":
    main()

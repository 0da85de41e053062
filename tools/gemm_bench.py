#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM generations on the codec's own shapes (through the C ABI test hook).
Prints TFLOP/s per backend: 2 = gen 1 (cta_group::1, direct stores), 3 = gen 2 single CTA + TMA store, 4 = gen 2 CTA pairs."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simwhisper_codec_b200 import _lib  # noqa: E402

SHAPES = [  # (name, M, N, K, out_bf16, act)
    ("qkv", 96000, 2304, 768, 1, 0),
    ("out_proj", 96000, 768, 768, 0, 0),
    ("fc1+gelu", 96000, 3072, 768, 1, 2),
    ("fc2", 96000, 768, 3072, 0, 0),
    ("pw1+gelu", 192000, 4096, 512, 1, 2),
    ("pw2", 192000, 512, 4096, 0, 0),
]


WARM = int(os.environ.get('GB_WARM', 3))
REPS = int(os.environ.get('GB_REPS', 10))


def main():
    lib = _lib.load()
    backends = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["2", "3", "4"])]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rows = []
    for name, M, N, K, obf, act in SHAPES:
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
        b = torch.randn(N, device="cuda", generator=g)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
        ref = None
        for be in backends:
            def run():
                _lib.check(lib.swc_test_gemm(be, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(b.data_ptr()),
                                             C.c_void_p(out.data_ptr()), obf, M, N, K, act, st), "gemm")
            for _ in range(WARM):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = REPS
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            o = out[:4096].float().clone()
            if ref is None:
                ref = o
                err = 0.0
            else:
                err = float((o - ref).abs().max())
            rows.append({"shape": name, "M": M, "N": N, "K": K, "backend": be, "ms": round(ms, 4),
                         "tflops": round(2.0 * M * N * K / ms / 1e9, 1), "max_abs_diff_vs_first": err})
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()

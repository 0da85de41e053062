#!/usr/bin/env python
"""Micro-benchmark of the attention kernels at the codec's shape (B x 1500 tokens x 12 heads x 64).
backends: 3 = bf16 tcgen05, 2 = three-product (bf16x3) tcgen05 behind an fp32 interface (the timing of that entry includes
the plane split / merge passes and two cudaMalloc: use the whole-model bench for its real cost), 1 / 0 = SIMT bf16 / fp32."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simwhisper_codec_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    backends = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["3"])]
    B, T, H = int(os.environ.get("AB_B", 64)), int(os.environ.get("AB_T", 1500)), 12
    reps = int(os.environ.get("AB_REPS", 5))
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g) * 0.7
    base[..., : H * 64] *= 0.125          # the packed q_proj carries the 1/sqrt(head_dim) scale
    lens = torch.full((B,), T, device="cuda", dtype=torch.int64)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = None
    for be in backends:
        dt = torch.float32 if be in (0, 2) else torch.bfloat16
        qkv = base.to(dt)
        out = torch.empty(B, T, H * 64, device="cuda", dtype=dt)

        def run():
            _lib.check(lib.swc_test_attention(be, C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(lens.data_ptr()),
                                              B, T, H, st), "attention")
        for _ in range(int(os.environ.get("AB_WARM", 2))):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        o = out.float().clone()
        err = 0.0 if ref is None else float((o - ref).abs().max())
        if ref is None:
            ref = o
        print(json.dumps({"backend": be, "B": B, "T": T, "ms": round(ms, 4), "tflops": round(4.0 * B * H * T * T * 64 / ms / 1e9, 1),
                          "max_abs_diff_vs_first": err}), flush=True)


if __name__ == "__main__":
    main()

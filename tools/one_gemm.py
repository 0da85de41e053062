#!/usr/bin/env python
"""Two launches of one GEMM shape through the C-ABI test hook (for `ncu -k regex:gemm_tc -s 1 -c 1`):
python tools/one_gemm.py M N K out_bf16 act backend   (backend: see tools/gemm_bench.py)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simwhisper_codec_b200 import _lib  # noqa: E402

lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M, N, K, obf, act, be = [int(x) for x in sys.argv[1:7]]
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
for _ in range(2):
    _lib.check(lib.swc_test_gemm(be, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(b.data_ptr()),
                                 C.c_void_p(out.data_ptr()), obf, M, N, K, act, st), "gemm")
torch.cuda.synchronize()
print("ok")

#!/usr/bin/env python
"""Phase timeline of one persistent CTA of the tcgen05 attention kernel (clock64 stamps of the first softmax warp of query
tile 0): A item decoded | per key block: B S ready, C P handed over | D last P V done | E item stored.
`--fine`: four more stamps per key block (they perturb the kernel by ~20 %).  Both query tiles are traced; the raw stamps go
to gpurun_out/attn_trace.npy when that directory exists."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simwhisper_codec_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    fn = lib.swc_debug_attn_trace
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_int]
    B, T, H = 64, 1500, 12
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g) * 0.7
    qkv[..., : H * 64] *= 0.125
    qkv = qkv.bfloat16()
    lens = torch.full((B,), T, device="cuda", dtype=torch.int64)
    out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        _lib.check(lib.swc_test_attention(3, C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(lens.data_ptr()),
                                          B, T, H, st), "attention")
    run()
    torch.cuda.synchronize()
    fine = "--fine" in sys.argv
    assert fn(3 if fine else 1, None, 0) == 0
    run()
    buf = np.zeros(3 * 4096, dtype=np.int64)
    assert fn(0, buf.ctypes.data, buf.size) == 0
    if os.path.isdir("gpurun_out"):
        np.save("gpurun_out/attn_trace.npy", buf)
    n_kt = 12
    k = 6 if fine else 2
    per = 1 + k * n_kt + 2
    t0, t1 = buf[:4096], buf[4096:8192]
    n = min((t0 > 0).sum(), (t1 > 0).sum()) // per
    b0 = t0[: n * per].reshape(n, per)[:, 1:1 + k * n_kt].reshape(n, n_kt, k)[:, :, 0]
    b1 = t1[: n * per].reshape(n, per)[:, 1:1 + k * n_kt].reshape(n, n_kt, k)[:, :, 0]
    print("tile 1 'S ready' minus tile 0 'S ready', per key block of item 3:", (b1[3] - b0[3]).tolist())
    print("median offset over items 2..:", int(np.median((b1 - b0)[2:])))
    t = buf[:4096]
    t = t[t > 0]
    # the default build stamps B (S ready) and C (P handed over) per key block; a library built with
    # -DSWC_ATTN_FINE_TRACE adds B1 (S in registers), B2 (row maximum known), B3 (exponentials done), B4 (previous P V done)
    items = len(t) // per
    t = t[: items * per].reshape(items, per)
    A, D, E = t[:, 0], t[:, -2], t[:, -1]
    blk = t[:, 1:1 + k * n_kt].reshape(items, n_kt, k)
    if not fine:      # B, C only: fill the intermediate stamps with B so that their differences read 0
        blk = np.stack([blk[:, :, 0]] * 5 + [blk[:, :, 1]], axis=2)
    B_, C_ = blk[:, :, 0], blk[:, :, 5]
    print(f"{items} items traced; cycles (median over items 2..):")
    sl = slice(2, None)
    med = lambda a: int(np.median(a[sl]))
    print("  item period (A -> next A)        ", med(np.diff(A)))
    print("  A -> B0 (first S ready)          ", med(B_[:, 0] - A))
    if fine:
        print("  B -> B1 (S: TMEM -> registers)   ", med(blk[:, :, 1] - blk[:, :, 0]))
        print("  B1 -> B2 (row max + exchange)    ", med(blk[:, :, 2] - blk[:, :, 1]))
        print("  B2 -> B3 (exponentials)          ", med(blk[:, :, 3] - blk[:, :, 2]))
        print("  B3 -> B4 (wait previous P V)     ", med(blk[:, 1:, 4] - blk[:, 1:, 3]))
        print("  B4 -> C (P -> TMEM, hand over)   ", med(blk[:, :, 5] - blk[:, :, 4]))
    else:
        print("  B_j -> C_j (softmax of a block)  ", med(C_ - B_))
    print("  C_j -> B_j+1 (wait for next S)   ", med(B_[:, 1:] - C_[:, :-1]))
    print("  C_last -> D (last P V)           ", med(D - C_[:, -1]))
    print("  D -> E (O read, normalise, store)", med(E - D))
    print("  E -> next A (decode next item)   ", med(A[1:] - E[:-1]))
    print("  per-block period B_j -> B_j+1    ", med(np.diff(B_, axis=1)))
    print("  block periods of one item        ", np.diff(B_[3]).tolist())


if __name__ == "__main__":
    main()

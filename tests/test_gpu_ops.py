"""Operator-level GPU tests through the C ABI: GEMM back ends and attention vs plain torch."""
import ctypes as C

import pytest
import torch

from simwhisper_codec_b200 import _lib

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _gemm(backend, A, W, bias, out_bf16, act=0):
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _lib.check(lib.swc_test_gemm(backend, _p(A), _p(W), _p(bias), _p(out), int(out_bf16), M, N, K, act, _stream()), "gemm")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 768, 768), (1000, 2304, 768), (257, 32, 512), (130, 656, 512),
                                   (4096, 4096, 512), (515, 512, 4096)])
def test_gemm_simt_fp32(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ W.double().T + b.double())
    out = _gemm(0, A, W, b, False)
    assert (out.double() - ref).abs().max().item() < 2e-4 * ref.abs().max().item()
    out_g = _gemm(0, A, W, b, False, act=1)
    assert (out_g.double() - torch.nn.functional.gelu(ref)).abs().max().item() < 2e-4 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (300, 768, 768), (1000, 2304, 768), (257, 32, 512),
                                   (130, 656, 512), (4096, 4096, 512), (515, 512, 4096), (20000, 3072, 768)])
@pytest.mark.parametrize("backend", [2, 3, 4])
def test_gemm_tcgen05_bf16(M, N, K, backend):
    """backend 2 = first-generation kernel, 3 = TMA-store epilogue, 4 = CTA-pair MMA (shapes the second generation
    does not take fall back to the first)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    ref = A.double() @ W.double().T + b.double()
    out = _gemm(backend, A, W, b, False)
    err = (out.double() - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), err       # fp32 accumulation of exact bf16 products
    out16 = _gemm(backend, A, W, b, True)
    assert (out16.double() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    # fused erf-GELU epilogue (tanh-form fit, 2.5e-5 + tanh.approx error) against exact GELU, bf16 output
    gel = _gemm(backend, A, W, b, True, act=2)
    gref = torch.nn.functional.gelu(ref)
    assert (gel.double() - gref).abs().max().item() < 1.5e-2 * max(1.0, gref.abs().max().item())
    gel32 = _gemm(2, A, W, b, False, act=2)
    assert (gel32.double() - gref).abs().max().item() < 2e-3 * max(1.0, gref.abs().max().item())
    # the SIMT kernel on the same bf16 operands must agree to fp32 rounding
    simt = _gemm(1, A, W, b, False)
    assert (simt.double() - out.double()).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())


def _attn_ref(qkv, lens, H):
    B, T, _ = qkv.shape
    q, k, v = (t.view(B, T, H, 64).transpose(1, 2).double() for t in qkv.float().split(H * 64, dim=-1))
    s = q @ k.transpose(-1, -2)
    ok = torch.arange(T, device=qkv.device)[None, :] < lens[:, None]
    s = s.masked_fill(~ok[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, H * 64)
    return o * ok[..., None]


@pytest.mark.parametrize("backend,dtype,tol", [(0, torch.float32, 2e-5), (1, torch.bfloat16, 2e-2), (2, torch.float32, 2e-5),
                                                (3, torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,T,lens", [(2, 100, [100, 37]), (3, 257, [257, 1, 130]), (1, 1500, [1500]), (2, 375, [64, 375]),
                                      (3, 1500, [1500, 129, 700])])
def test_attention(backend, dtype, tol, B, T, lens):
    """backend 3 = tcgen05/TMEM flash attention (the product path in bf16 mode), 2 = the three-product tcgen05 kernel of the
    bf16x3 mode behind an fp32 interface (held to the fp32 kernel's tolerance), 0 / 1 = SIMT fp32 / bf16."""
    lib = _lib.load()
    H = 12
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = (torch.randn(B, T, 3 * H * 64, device="cuda", generator=g) * 0.7).to(dtype)
    qkv[..., : H * 64] *= 0.125
    lens_t = torch.tensor(lens, device="cuda", dtype=torch.int64)
    out = torch.full((B, T, H * 64), float("nan"), device="cuda", dtype=dtype)
    _lib.check(lib.swc_test_attention(backend, _p(qkv), _p(out), _p(lens_t), B, T, H, _stream()), "attention")
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, lens_t, H)
    ok = torch.arange(T, device="cuda")[None, :] < lens_t[:, None]
    diff = ((out.double() - ref) * ok[..., None]).abs().max().item()
    assert diff < tol, diff
    assert torch.isfinite(out).all()          # padded query rows are written (zeros), never left as garbage


@pytest.mark.parametrize("backend", [2, 3])
def test_attention_growing_maximum(backend):
    """Scores whose row maximum keeps growing along the keys (ramp) force the online-softmax rescale path
    (the tcgen05 kernel rescales lazily, only when the maximum grew by more than 2^8)."""
    lib = _lib.load()
    B, T, H = 2, 640, 12
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g) * 0.5
    ramp = torch.linspace(0.0, 6.0, T, device="cuda")
    qkv[..., H * 64: 2 * H * 64] += ramp[None, :, None] * 0.3          # keys drift along time
    qkv[..., : H * 64] = qkv[..., : H * 64].abs() * 0.4                # positive queries -> scores grow with the key index
    dtype, tol = (torch.float32, 2e-5) if backend == 2 else (torch.bfloat16, 2e-2)
    qkv = qkv.to(dtype)
    lens_t = torch.tensor([T, 333], device="cuda", dtype=torch.int64)
    out = torch.full((B, T, H * 64), float("nan"), device="cuda", dtype=dtype)
    _lib.check(lib.swc_test_attention(backend, _p(qkv), _p(out), _p(lens_t), B, T, H, _stream()), "attention")
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, lens_t, H)
    ok = torch.arange(T, device="cuda")[None, :] < lens_t[:, None]
    diff = ((out.double() - ref) * ok[..., None]).abs().max().item()
    assert diff < tol, diff


@pytest.mark.parametrize("backend", [2, 3])
def test_attention_many_items_mixed_lengths(backend):
    """More work items than SMs with every kind of item mixed on one persistent CTA: full tiles, tiles whose second
    128-query half is padding, fully padded (dead) tiles, one-key sequences and empty sequences."""
    lib = _lib.load()
    B, T, H = 40, 1500, 12
    g = torch.Generator().manual_seed(11)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[:8] = torch.tensor([1500, 1, 0, 128, 129, 256, 257, 1407])
    qkv = torch.randn(B, T, 3 * H * 64, generator=g) * 0.7
    qkv[..., : H * 64] *= 0.125
    dtype, tol = (torch.float32, 2e-5) if backend == 2 else (torch.bfloat16, 2e-2)
    qkv = qkv.to(dtype).cuda()
    lens_t = lens.to(device="cuda", dtype=torch.int64)
    out = torch.full((B, T, H * 64), float("nan"), device="cuda", dtype=dtype)
    _lib.check(lib.swc_test_attention(backend, _p(qkv), _p(out), _p(lens_t), B, T, H, _stream()), "attention")
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    ok = torch.arange(T, device="cuda")[None, :] < lens_t[:, None]
    assert float(out.float().abs().amax(dim=-1)[~ok].max()) == 0.0           # padded query rows are zeros
    for b in (0, 1, 3, 4, 6, 7, 20, 39):                                      # reference on a subset (memory)
        ref = _attn_ref(qkv[b:b + 1], lens_t[b:b + 1], H)
        diff = ((out[b:b + 1].double() - ref) * ok[b:b + 1, :, None]).abs().max().item()
        assert diff < tol, (b, int(lens[b]), diff)

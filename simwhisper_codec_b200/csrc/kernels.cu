// Bandwidth-bound stages of the codec: LayerNorm, depthwise-conv+LayerNorm, anti-aliased Snake,
// FSQ, layout conversion, log-mel pre/post passes, iSTFT overlap-add.  All are HBM-streaming
// kernels: coalesced along the channel (contiguous) dimension, 16/32-byte vector accesses, fp32 math.
#include <algorithm>
#include <cstdlib>

#include <type_traits>

#include "kernels.cuh"

namespace swc {

// ================================================================================================
// LayerNorm: one warp per row; each lane owns NCH chunks of 8 contiguous channels.
// ================================================================================================
template <typename TO, int NCH>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, const float* __restrict__ delta,
                                                        float* __restrict__ h_out, TO* __restrict__ out,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int nb,
                                                        int t_in, int t_out, const long long* __restrict__ lens) {
  constexpr int C = NCH * 256;
  constexpr int RW = 2;                                               // rows per warp: both rows' loads are in flight together
  constexpr bool kPlanes = std::is_same<TO, bf16_planes>::value;     // rows of (hi | lo) bf16 planes, C columns each
  using TE = typename std::conditional<kPlanes, bf16, TO>::type;
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RW;
  const long long n_rows = (long long)nb * t_out;
  if (row0 >= n_rows) return;
  float v[RW][NCH][8];
  bool exists[RW], in_range[RW], live[RW];
  long long irow[RW];
  // the length is fetched together with the rows, not before them: a row load that waits for lens[b] pays the L2 latency
  // twice (rows beyond the length are then read for nothing; variable-length batches run the packed path instead)
  long long len_b[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const long long row = row0 + r;
    exists[r] = row < n_rows;
    const int b = (int)((unsigned)(exists[r] ? row : row0) / (unsigned)t_out);     // nb * t_out < 2^31 (checked by the launcher)
    const int t = (int)((exists[r] ? row : row0) - (long long)b * t_out);
    in_range[r] = exists[r] && t < t_in;
    len_b[r] = (in_range[r] && lens) ? lens[b] : (long long)t_in;
    irow[r] = ((long long)b * t_in + t) * C;
    if (in_range[r]) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int c0 = (c * 32 + lane) * 8;
        load8(in + irow[r] + c0, v[r][c]);
        if (delta) {      // fused residual add: h <- h + delta (the GEMM that produced delta has no residual epilogue)
          float dv[8];
          load8(delta + irow[r] + c0, dv);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[r][c][j] += dv[j];
          if (h_out) store8(h_out + irow[r] + c0, v[r][c]);
        }
      }
    }
    live[r] = in_range[r] && (long long)t < len_b[r];
  }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    if (!exists[r]) continue;
    TE* o = reinterpret_cast<TE*>(out) + (row0 + r) * (kPlanes ? 2 * C : C);
    if (!live[r]) {
      float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        store8(o + (c * 32 + lane) * 8, z);
        if constexpr (kPlanes) store8(o + C + (c * 32 + lane) * 8, z);
      }
      continue;
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[r][c][j];
    const float mean = warp_sum(sum) * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) { float dlt = v[r][c][j] - mean; sq = fmaf(dlt, dlt, sq); }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / C) + eps);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int c0 = (c * 32 + lane) * 8;
      float g[8], bt[8], res[8];
      load8(gamma + c0, g);
      load8(beta + c0, bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) res[j] = (v[r][c][j] - mean) * rstd * g[j] + bt[j];
      if constexpr (kPlanes) store8_planes(o + c0, C, res);
      else store8(o + c0, res);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Streaming form for the common case (contiguous rows, no length mask, no fused add): a producer warp streams 8-row chunks
// into a 4-stage shared-memory ring with bulk async copies (cp.async.bulk + mbarrier transaction bytes), eight consumer warps
// normalise one row each straight from shared memory.  The loads are decoupled from the arithmetic: up to 4 x 24 KB per CTA
// (two CTAs per SM) are in flight whatever the consumers are doing.  Same lane -> channel mapping and the same order of
// additions as layernorm_kernel: results are bit-identical.
// ------------------------------------------------------------------------------------------------
namespace lnstream {
constexpr int kRows = 8, kStages = 4, kThreads = 32 * (kRows + 1);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
}  // namespace lnstream

template <typename TO, int NCH>
__global__ void __launch_bounds__(lnstream::kThreads, 2) layernorm_stream_kernel(const float* __restrict__ in, TO* __restrict__ out,
                                                                                const float* __restrict__ gamma,
                                                                                const float* __restrict__ beta, float eps, long long n_rows) {
  using namespace lnstream;
  constexpr int C = NCH * 256;
  constexpr bool kPlanes = std::is_same<TO, bf16_planes>::value;
  using TE = typename std::conditional<kPlanes, bf16, TO>::type;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  float* ring = reinterpret_cast<float*>(ln_smem);                                  // [kStages][kRows][C]
  uint64_t* full = reinterpret_cast<uint64_t*>(ln_smem + (size_t)kStages * kRows * C * 4);
  uint64_t* empty = full + kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kRows); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long n_chunks = (n_rows + kRows - 1) / kRows;
  if (warp == kRows) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
        const uint32_t st = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        const long long r0 = c * kRows;
        const uint32_t bytes = (uint32_t)(min((long long)kRows, n_rows - r0) * C * 4);
        mbar_expect_tx(&full[st], bytes);
        bulk_load(ring + (size_t)st * kRows * C, in + r0 * C, bytes, &full[st]);
      }
    }
    return;
  }
  float g[NCH][8], bt[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    load8(gamma + (c * 32 + lane) * 8, g[c]);
    load8(beta + (c * 32 + lane) * 8, bt[c]);
  }
  uint32_t it = 0;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
    const uint32_t st = it % kStages, ph = (it / kStages) & 1;
    const long long row = c * kRows + warp;
    mbar_wait(&full[st], ph);
    if (row < n_rows) {
      const float* src = ring + ((size_t)st * kRows + warp) * C;
      float v[NCH][8];
#pragma unroll
      for (int k = 0; k < NCH; ++k) load8(src + (k * 32 + lane) * 8, v[k]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[k][j];
      const float mean = warp_sum(sum) * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) { float dlt = v[k][j] - mean; sq = fmaf(dlt, dlt, sq); }
      const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / C) + eps);
      TE* o = reinterpret_cast<TE*>(out) + row * (kPlanes ? 2 * C : C);
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c0 = (k * 32 + lane) * 8;
        float res[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) res[j] = (v[k][j] - mean) * rstd * g[k][j] + bt[k][j];
        if constexpr (kPlanes) store8_planes(o + c0, C, res);
        else store8(o + c0, res);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
}

template <typename TO, int NCH>
static int layernorm_stream_launch(const float* in, void* out, const float* g, const float* b, float eps, long long rows, cudaStream_t s) {
  constexpr int C = NCH * 256;
  const int smem = lnstream::kStages * lnstream::kRows * C * 4 + 2 * lnstream::kStages * 8;
  auto kern = layernorm_stream_kernel<TO, NCH>;
  SWC_TRY(ensure_dynamic_smem((const void*)kern, smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long chunks = (rows + lnstream::kRows - 1) / lnstream::kRows;
  const int grid = (int)std::min<long long>(chunks, 2ll * sms);
  ProfScope ps(KC_LAYERNORM, s);
  kern<<<grid, lnstream::kThreads, smem, s>>>(in, (TO*)out, g, b, eps, rows);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename TO>
static int layernorm_t(const float* in, const float* delta, float* h_out, void* out, const float* g, const float* b, float eps,
                       int nb, int t_in, int t_out, int C, const long long* lens, cudaStream_t s) {
  const long long rows = (long long)nb * t_out;
  SWC_REQUIRE(rows < (1ll << 31), "layernorm: too many rows (%lld)", rows);
  const int warps = 8;
  dim3 grid((unsigned)ceil_div_ll(rows, warps * 2));      // two rows per warp
  ProfScope ps(KC_LAYERNORM, s);
  if (C == 768) layernorm_kernel<TO, 3><<<grid, warps * 32, 0, s>>>(in, delta, h_out, (TO*)out, g, b, eps, nb, t_in, t_out, lens);
  else if (C == 512) layernorm_kernel<TO, 2><<<grid, warps * 32, 0, s>>>(in, delta, h_out, (TO*)out, g, b, eps, nb, t_in, t_out, lens);
  else { set_error("layernorm: unsupported width %d", C); return -1; }
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int layernorm(const float* in, const float* delta, float* h_out, void* out, int out_type, const float* gamma,
              const float* beta, float eps, int nb, int t_in, int t_out, int C, const long long* lens, cudaStream_t s) {
  static const bool stream_ok = [] { const char* e = getenv("SWC_LN_STREAM"); return !(e && e[0] == '0'); }();
  if (stream_ok && !delta && !lens && t_in == t_out && ((uintptr_t)in & 15) == 0 && (long long)nb * t_in >= 4096 && (C == 768 || C == 512)) {
    const long long rows = (long long)nb * t_in;
    if (C == 768) {
      if (out_type == 0) return layernorm_stream_launch<float, 3>(in, out, gamma, beta, eps, rows, s);
      if (out_type == 2) return layernorm_stream_launch<bf16_planes, 3>(in, out, gamma, beta, eps, rows, s);
      return layernorm_stream_launch<bf16, 3>(in, out, gamma, beta, eps, rows, s);
    }
    if (out_type == 0) return layernorm_stream_launch<float, 2>(in, out, gamma, beta, eps, rows, s);
    if (out_type == 2) return layernorm_stream_launch<bf16_planes, 2>(in, out, gamma, beta, eps, rows, s);
    return layernorm_stream_launch<bf16, 2>(in, out, gamma, beta, eps, rows, s);
  }
  if (out_type == 0) return layernorm_t<float>(in, delta, h_out, out, gamma, beta, eps, nb, t_in, t_out, C, lens, s);
  if (out_type == 2) return layernorm_t<bf16_planes>(in, delta, h_out, out, gamma, beta, eps, nb, t_in, t_out, C, lens, s);
  return layernorm_t<bf16>(in, delta, h_out, out, gamma, beta, eps, nb, t_in, t_out, C, lens, s);
}

// ================================================================================================
// depthwise conv k7 + bias + LayerNorm (Vocos ConvNeXt block, reference modules.py:1232-1240)
//
// HBM-streaming: every element of x (and delta) is read once (+ 6 halo rows per strip), x + delta and the
// normalised row are written once.  One block = 128 threads = the 512 channels (4 per thread) of one time strip
// of one item; it slides along time with a 14-row register window of s = x + delta, producing 8 output rows per
// step: 16 independent 16-byte loads in flight per thread, the 7-tap convolution entirely in registers, and the
// LayerNorm statistics (two-pass, as torch) reduced across the 4 warps through shared memory once per 8 rows.
// ================================================================================================
constexpr int kDwRows = 8;       // output rows per step
constexpr int kDwThreads = 128;  // 4 channels per thread, C = 512

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// Sums of 8 per-lane values over the warp in 9 shuffles instead of 40: each exchange halves the number of values a lane
// carries (lane bit 4 / 3 / 2 picks which half it keeps), so after three exchanges a lane owns one row, and two more
// exchanges finish that row.  Lane l ends with the total of row l >> 2; the lane pairs and their order are those of
// warp_sum, so every total is bit-identical to it.
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float a[4], b[2];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float recv = __shfl_xor_sync(0xffffffffu, b4 ? v[k] : v[k + 4], 16);
    a[k] = (b4 ? v[k + 4] : v[k]) + recv;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float recv = __shfl_xor_sync(0xffffffffu, b3 ? a[k] : a[k + 2], 8);
    b[k] = (b3 ? a[k + 2] : a[k]) + recv;
  }
  float c = (b2 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, b2 ? b[0] : b[1], 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}
// four channels as two packed fp32 pairs: fma / mul / add / sub .f32x2 (sm_100) handle two lanes per issued instruction
struct F4P { unsigned long long lo, hi; };
__device__ __forceinline__ unsigned long long pk2f(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2f(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ F4P f4p(float4 v) { return F4P{pk2f(v.x, v.y), pk2f(v.z, v.w)}; }
__device__ __forceinline__ F4P f4p(float v) { const unsigned long long p = pk2f(v, v); return F4P{p, p}; }
__device__ __forceinline__ float4 f4(F4P v) { float4 r; upk2f(v.lo, r.x, r.y); upk2f(v.hi, r.z, r.w); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ F4P fma(F4P a, F4P b, F4P c) { return F4P{fma2(a.lo, b.lo, c.lo), fma2(a.hi, b.hi, c.hi)}; }
__device__ __forceinline__ F4P mul(F4P a, F4P b) { return F4P{mul2(a.lo, b.lo), mul2(a.hi, b.hi)}; }
__device__ __forceinline__ F4P sub(F4P a, F4P b) { return F4P{sub2(a.lo, b.lo), sub2(a.hi, b.hi)}; }

template <typename TO>
__device__ __forceinline__ float dw_rstd(float var_eps) {
  if constexpr (sizeof(TO) == 2) return rsqrtf(var_eps);      // bf16 output: MUFU.RSQ (1 ulp) instead of sqrt + divide
  else return 1.0f / sqrtf(var_eps);
}

// TAB = DenseRows: item b owns rows [b T, b T + T).  TAB = RaggedTable: item b owns rows [off[b], off[b] + len[b]) of a
// packed buffer (items of different lengths back to back, Vocos over valid frames + halo only); rows outside an item are
// the convolution's zero padding in both layouts, so neighbouring items never see each other.
struct DenseRows {};
template <typename TO, bool HAS_DELTA, typename TAB>
__global__ void __launch_bounds__(kDwThreads, 4) dwconv7_ln_kernel(const float* __restrict__ x, const float* __restrict__ delta,
                                                                   float* __restrict__ x_out, const float* __restrict__ w,
                                                                   const float* __restrict__ bias,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float eps,
                                                                   TO* __restrict__ out, int T_dense, int strip,
                                                                   const __grid_constant__ TAB tab) {
  constexpr int C = 512, R = kDwRows;
  constexpr bool kPlanes = std::is_same<TO, bf16_planes>::value;     // rows of (hi | lo) bf16 planes (bf16x3 mode: exact rstd)
  using TE = typename std::conditional<kPlanes, bf16, TO>::type;
  __shared__ __align__(16) float red_sum[R][4];
  __shared__ __align__(16) float red_sq[R][4];
  __shared__ __align__(16) float sw[7 * C];            // taps: read back per step instead of living in 28 registers
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = tid * 4;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * strip;
  int T;
  long long row0;
  if constexpr (std::is_same<TAB, RaggedTable>::value) {
    T = tab.len[b];
    row0 = tab.off[b];
    if (t0 >= T) return;                  // strips beyond this item's rows (the grid spans the longest item)
  } else {
    T = T_dense;
    row0 = (long long)b * T;
  }
  const int t_end = min(T, t0 + strip);
  const float* xb = x + row0 * C + c0;
  const float* db = HAS_DELTA ? delta + row0 * C + c0 : nullptr;
  float* xo = HAS_DELTA ? x_out + row0 * C + c0 : nullptr;
  TE* ob = reinterpret_cast<TE*>(out) + row0 * (kPlanes ? 2 * C : C) + c0;

#pragma unroll
  for (int k = 0; k < 7; ++k) *reinterpret_cast<float4*>(sw + k * C + c0) = *reinterpret_cast<const float4*>(w + k * C + c0);
  const float4 bs = *reinterpret_cast<const float4*>(bias + c0);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // (each thread reads back only the taps it wrote: no barrier needed)

  // window[j] holds s(row base - 3 + j); rows outside [0, T) are the convolution's zero padding
  F4P win[R + 6];
  auto load_row = [&](int t, bool own) -> float4 {
    if (t < 0 || t >= T) return zero4;
    float4 v = *reinterpret_cast<const float4*>(xb + (long long)t * C);
    if (HAS_DELTA) {
      v = f4_add(v, *reinterpret_cast<const float4*>(db + (long long)t * C));
      if (own) *reinterpret_cast<float4*>(xo + (long long)t * C) = v;   // the updated residual stream, written by the strip that owns the row
    }
    return v;
  };
#pragma unroll
  for (int j = 0; j < 6; ++j) win[j] = f4p(load_row(t0 - 3 + j, j >= 3));

  // rows base+3 .. base+10 of the coming step (the last three of the strip's final step belong to the next strip);
  // the loads of step n+1 are issued before the LayerNorm phase of step n so that their latency hides behind it
  float4 xv[R], dv[R];
  auto fetch = [&](int base) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int t = base + 3 + j;
      const bool ok = t < T;
      xv[j] = ok ? *reinterpret_cast<const float4*>(xb + (long long)t * C) : zero4;
      if (HAS_DELTA) dv[j] = ok ? *reinterpret_cast<const float4*>(db + (long long)t * C) : zero4;
    }
  };
  fetch(t0);
  const F4P bsp = f4p(bs);
  const F4P gmp = f4p(*reinterpret_cast<const float4*>(gamma + c0));
  const F4P btp = f4p(*reinterpret_cast<const float4*>(beta + c0));
  for (int base = t0; base < t_end; base += R) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int t = base + 3 + j;
      float4 v = xv[j];
      if (HAS_DELTA) {
        v = f4_add(v, dv[j]);
        if (t < t_end) *reinterpret_cast<float4*>(xo + (long long)t * C) = v;
      }
      win[6 + j] = f4p(v);
    }
    // conv + bias for rows base .. base+7 (same fused multiply-adds in the same order as the scalar form, two per instruction)
    F4P y[R];
    float psum[R];
#pragma unroll
    for (int r = 0; r < R; ++r) y[r] = bsp;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const F4P wk = f4p(*reinterpret_cast<const float4*>(sw + k * C + c0));
#pragma unroll
      for (int r = 0; r < R; ++r) y[r] = fma(win[r + k], wk, y[r]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) win[j] = win[j + R];
    if (base + R < t_end) fetch(base + R);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 v = f4(y[r]);
      psum[r] = (v.x + v.y) + (v.z + v.w);
    }
    {
      const float tot = warp_sum8(psum, lane);          // lane l: row l >> 2
      if ((lane & 3) == 0) red_sum[lane >> 2][warp] = tot;
    }
    __syncthreads();
    float mean[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 p = *reinterpret_cast<const float4*>(red_sum[r]);
      mean[r] = ((p.x + p.y) + (p.z + p.w)) * (1.0f / C);
      y[r] = sub(y[r], f4p(mean[r]));                   // centred from here on
      float a, b2;
      upk2f(fma2(y[r].lo, y[r].lo, mul2(y[r].hi, y[r].hi)), a, b2);   // (dx^2 + dz^2, dy^2 + dw^2)
      psum[r] = a + b2;
    }
    {
      const float tot = warp_sum8(psum, lane);
      if ((lane & 3) == 0) red_sq[lane >> 2][warp] = tot;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = base + r;
      if (t < t_end) {
        const float4 p = *reinterpret_cast<const float4*>(red_sq[r]);
        const float rstd = dw_rstd<TO>(((p.x + p.y) + (p.z + p.w)) * (1.0f / C) + eps);
        const float4 res = f4(fma(mul(y[r], f4p(rstd)), gmp, btp));
        if constexpr (kPlanes) store4_planes(ob + (long long)t * 2 * C, C, res);
        else store4(ob + (long long)t * C, res);
      }
    }
  }
}

namespace {
template <typename TAB>
int dwconv7_ln_launch(const float* x, const float* delta, float* x_out, const float* w7c, const float* bias, const float* gamma,
                      const float* beta, float eps, void* out, int out_type, int nb, int T, const TAB& tab, cudaStream_t s) {
  ProfScope ps(KC_DWCONV_LN, s);
  // strips of 64 rows (9 % halo re-reads, served by the L2) when that gives every SM many blocks, else 32
  const int strip = ((long long)nb * ceil_div(T, 64) >= 16 * 148) ? 64 : 32;
  dim3 grid(ceil_div(T, strip), nb);
  if (out_type == 2) {
    if (delta) dwconv7_ln_kernel<bf16_planes, true, TAB><<<grid, kDwThreads, 0, s>>>(x, delta, x_out, w7c, bias, gamma, beta, eps, (bf16_planes*)out, T, strip, tab);
    else dwconv7_ln_kernel<bf16_planes, false, TAB><<<grid, kDwThreads, 0, s>>>(x, nullptr, nullptr, w7c, bias, gamma, beta, eps, (bf16_planes*)out, T, strip, tab);
  } else if (out_type == 0) {
    if (delta) dwconv7_ln_kernel<float, true, TAB><<<grid, kDwThreads, 0, s>>>(x, delta, x_out, w7c, bias, gamma, beta, eps, (float*)out, T, strip, tab);
    else dwconv7_ln_kernel<float, false, TAB><<<grid, kDwThreads, 0, s>>>(x, nullptr, nullptr, w7c, bias, gamma, beta, eps, (float*)out, T, strip, tab);
  } else {
    if (delta) dwconv7_ln_kernel<bf16, true, TAB><<<grid, kDwThreads, 0, s>>>(x, delta, x_out, w7c, bias, gamma, beta, eps, (bf16*)out, T, strip, tab);
    else dwconv7_ln_kernel<bf16, false, TAB><<<grid, kDwThreads, 0, s>>>(x, nullptr, nullptr, w7c, bias, gamma, beta, eps, (bf16*)out, T, strip, tab);
  }
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

int dwconv7_ln(const float* x, const float* delta, float* x_out, const float* w7c, const float* bias, const float* gamma,
               const float* beta, float eps, void* out, int out_type, int nb, int T, int C, cudaStream_t s) {
  SWC_REQUIRE(C == 512, "dwconv7_ln: only C=512 is built (got %d)", C);
  SWC_REQUIRE(!delta || (x_out && x_out != x), "dwconv7_ln: fused residual needs a distinct output stream buffer");
  return dwconv7_ln_launch(x, delta, x_out, w7c, bias, gamma, beta, eps, out, out_type, nb, T, DenseRows{}, s);
}

int dwconv7_ln_ragged(const float* x, const float* w7c, const float* bias, const float* gamma, const float* beta, float eps,
                      void* out, int out_type, const RaggedTable& tab, int C, cudaStream_t s) {
  SWC_REQUIRE(C == 512, "dwconv7_ln_ragged: only C=512 is built (got %d)", C);
  SWC_REQUIRE(tab.nb > 0 && tab.nb <= kMaxRagged && tab.t_max > 0, "dwconv7_ln_ragged: bad table");
  return dwconv7_ln_launch(x, nullptr, nullptr, w7c, bias, gamma, beta, eps, out, out_type, tab.nb, tab.t_max, tab, s);
}

// ================================================================================================
// anti-aliased SnakeBeta (reference alias_free_torch/act.py:23-27, resample.py:25-33, filter.py:83-92,
// activations.py:107-119; closed form in SURVEY.md appendix A4)
//   u[2q]   = 2 * sum_a x[cl(q-3+a)] f[11-2a]      u[2q+1] = 2 * sum_a x[cl(q-2+a)] f[10-2a]   a=0..5
//   v[m]    = u[m] + sin^2(u[m] e^alpha) / (e^beta + 1e-9)
//   y[t]    = sum_{k<12} v[cl2(2t+k-5)] f[k]
// one thread = one channel, walking TCH consecutive frames with a 12-deep sliding window of v.
// ================================================================================================

template <typename TI, typename TO>
__global__ void __launch_bounds__(128) aa_snake_kernel(const TI* __restrict__ in, TO* __restrict__ out,
                                                       const float* __restrict__ taps_up,
                                                       const float* __restrict__ taps_dn,
                                                       const float* __restrict__ alpha_log,
                                                       const float* __restrict__ beta_log, int T, int C, int chunk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * chunk;
  if (c >= C) return;
  float f[12], fd[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) { f[i] = __ldg(taps_up + i); fd[i] = __ldg(taps_dn + i); }
  const float ea = expf(alpha_log[c]);
  const float inv_b = 1.0f / (expf(beta_log[c]) + 1e-9f);
  const TI* x = in + (long long)b * T * C + c;
  TO* y = out + (long long)b * T * C + c;
  const int T2 = 2 * T;

  auto snake = [&](float u) -> float {
    float sn;
    if constexpr (sizeof(TO) == 2) {
      // bf16 output: two-constant Cody-Waite reduction to [-pi, pi] + MUFU.SIN (abs error ~2^-21, far below bf16 rounding)
      const float a = u * ea;
      const float k = rintf(a * 0.15915494309189535f);
      const float r = fmaf(k, 1.7484555e-7f, fmaf(k, -6.2831855f, a));   // a - k*2pi with 2pi = 6.2831855f - 1.7484555e-7f
      sn = __sinf(r);
    } else {
      sn = sinf(u * ea);
    }
    return u + inv_b * (sn * sn);
  };
  auto calc_v = [&](int m) -> float {
    m = min(max(m, 0), T2 - 1);
    const int q = m >> 1, odd = m & 1;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      int j = min(max(q - 3 + odd + a, 0), T - 1);
      float xv = to_f32<TI>(x[(long long)j * C]);
      float tap = odd ? f[10 - 2 * a] : f[11 - 2 * a];
      acc = fmaf(xv, tap, acc);
    }
    return snake(2.0f * acc);
  };

  float vw[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) vw[k] = calc_v(2 * t0 + k - 5);
  const int t1 = min(t0 + chunk, T);
  // The two new window entries of a step, v[2s+5] (odd) and v[2s+6] (even) with s = t + 1, read the same six inputs
  // x[cl(s) .. cl(s+5)]: a six-deep register window of x slides one frame per step (one load per frame instead of 12;
  // same taps in the same order as calc_v, so the result is bit-identical).
  float xs[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) xs[a] = to_f32<TI>(x[(long long)min(t0 + 1 + a, T - 1) * C]);
  for (int t = t0; t < t1; ++t) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc = fmaf(vw[k], fd[k], acc);
    y[(long long)t * C] = from_f32<TO>(acc);
    if (t + 1 < t1) {
      const float x_next = to_f32<TI>(x[(long long)min(t + 7, T - 1) * C]);     // x[cl(s + 6)], for the next step
#pragma unroll
      for (int k = 0; k < 10; ++k) vw[k] = vw[k + 2];
      const int s = t + 1;
      float ao = 0.f, ae = 0.f;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        ao = fmaf(xs[a], f[10 - 2 * a], ao);
        ae = fmaf(xs[a], f[11 - 2 * a], ae);
      }
      // positions past the end replicate v[2T-1] (the newest valid entry is then already in the window)
      vw[10] = (2 * s + 5 <= T2 - 1) ? snake(2.0f * ao) : vw[9];
      vw[11] = (2 * s + 6 <= T2 - 1) ? snake(2.0f * ae) : vw[10];
#pragma unroll
      for (int a = 0; a < 5; ++a) xs[a] = xs[a + 1];
      xs[5] = x_next;
    }
  }
}

int aa_snake(const void* in, int in_type, void* out, int out_type, const float* taps_up, const float* taps_dn,
             const float* alpha_log, const float* beta_log, int nb, int T, int C, cudaStream_t s) {
  // frames per thread: the 12-entry window fill costs as much as six frames, so 64-frame chunks when the grid stays large
  const int chunk = ((long long)nb * ceil_div(T, 64) * ceil_div(C, 128) >= 16 * 148) ? 64 : 32;
  dim3 grid(ceil_div(C, 128), ceil_div(T, chunk), nb);
  ProfScope ps(KC_SNAKE, s);
  if (in_type == 0 && out_type == 0) aa_snake_kernel<float, float><<<grid, 128, 0, s>>>((const float*)in, (float*)out, taps_up, taps_dn, alpha_log, beta_log, T, C, chunk);
  else if (in_type == 0 && out_type == 1) aa_snake_kernel<float, bf16><<<grid, 128, 0, s>>>((const float*)in, (bf16*)out, taps_up, taps_dn, alpha_log, beta_log, T, C, chunk);
  else if (in_type == 1 && out_type == 1) aa_snake_kernel<bf16, bf16><<<grid, 128, 0, s>>>((const bf16*)in, (bf16*)out, taps_up, taps_dn, alpha_log, beta_log, T, C, chunk);
  else { set_error("aa_snake: unsupported types %d->%d", in_type, out_type); return -1; }
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
// FSQ stand-alone kernels (reference quantizer.py:181-224, 273-317)
// ================================================================================================
__global__ void fsq_encode_cf_kernel(const float* __restrict__ z, const long long* __restrict__ lens, int nb, int T,
                                     FsqConst c, float* zq_cf, int* codes, float* zq_cl) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over (g, b, t)
  const long long total = 8LL * nb * T;
  if (i >= total) return;
  const int t = (int)(i % T);
  const int b = (int)((i / T) % nb);
  const int g = (int)(i / ((long long)T * nb));
  float x[4], dq[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) x[d] = z[((long long)b * 32 + g * 4 + d) * T + t];
  int idx = fsq_quantize4(c, x, dq);
  if ((long long)t >= lens[b]) { idx = 0; dq[0] = dq[1] = dq[2] = dq[3] = 0.f; }
  if (codes) codes[i] = idx;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    if (zq_cf) zq_cf[((long long)b * 32 + g * 4 + d) * T + t] = dq[d];
    if (zq_cl) zq_cl[((long long)b * T + t) * 32 + g * 4 + d] = dq[d];
  }
}

int fsq_encode_cf(const float* latent_cf, const long long* lens, int nb, int T, const FsqConst& c, float* zq_cf,
                  int* codes, float* zq_cl, cudaStream_t s) {
  const long long total = 8LL * nb * T;
  ProfScope ps(KC_MISC, s);
  fsq_encode_cf_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, s>>>(latent_cf, lens, nb, T, c, zq_cf, codes, zq_cl);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename TI>
__global__ void fsq_decode_kernel(const TI* __restrict__ codes, const long long* __restrict__ lens, int nb, int T,
                                  FsqConst c, float* zq_cf, float* zq_cl) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = 8LL * nb * T;
  if (i >= total) return;
  const int t = (int)(i % T);
  const int b = (int)((i / T) % nb);
  const int g = (int)(i / ((long long)T * nb));
  const long long idx = (long long)codes[i];
  const bool valid = (long long)t < lens[b];
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    // python floor-div/mod semantics coincide with C for the non-negative indices of a codebook
    const long long lvl = (idx / c.base[d]) % c.levels[d];
    float dq = ((float)lvl - c.half[d]) / c.half[d];
    if (!valid) dq = 0.f;
    if (zq_cf) zq_cf[((long long)b * 32 + g * 4 + d) * T + t] = dq;
    if (zq_cl) zq_cl[((long long)b * T + t) * 32 + g * 4 + d] = dq;
  }
}

int fsq_decode(const void* codes, int codes_i64, const long long* lens, int nb, int T, const FsqConst& c,
               float* zq_cf, float* zq_cl, cudaStream_t s) {
  const long long total = 8LL * nb * T;
  const unsigned grid = (unsigned)ceil_div_ll(total, 256);
  ProfScope ps(KC_MISC, s);
  if (codes_i64) fsq_decode_kernel<long long><<<grid, 256, 0, s>>>((const long long*)codes, lens, nb, T, c, zq_cf, zq_cl);
  else fsq_decode_kernel<int><<<grid, 256, 0, s>>>((const int*)codes, lens, nb, T, c, zq_cf, zq_cl);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
// layout conversion through a 32x33 shared tile
// ================================================================================================
template <typename TO>
__global__ void cf_to_cl_kernel(const float* __restrict__ in, TO* __restrict__ out, int C, int T, int t_rows, int c_pitch) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? in[((long long)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < t_rows && c < c_pitch) out[((long long)b * t_rows + t) * c_pitch + c] = from_f32<TO>(tile[tx][i]);
  }
}

int cf_to_cl(const float* in, void* out, int out_type, int nb, int C, int T, int t_rows, int c_pitch, cudaStream_t s) {
  dim3 grid(ceil_div(t_rows, 32), ceil_div(c_pitch, 32), nb), block(32, 8);
  ProfScope ps(KC_MISC, s);
  if (out_type == 0) cf_to_cl_kernel<float><<<grid, block, 0, s>>>(in, (float*)out, C, T, t_rows, c_pitch);
  else cf_to_cl_kernel<bf16><<<grid, block, 0, s>>>(in, (bf16*)out, C, T, t_rows, c_pitch);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename TI>
__global__ void cl_to_cf_kernel(const TI* __restrict__ in, float* __restrict__ out, int C, int T,
                                long long in_batch_stride, int c_pitch, const long long* __restrict__ lens) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t_valid = lens ? (int)(lens[b] < 0 ? 0 : (lens[b] > T ? T : lens[b])) : T;    // rows >= lens[b] read as zero
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    tile[i][tx] = (t < t_valid && c < C) ? to_f32<TI>(in[(long long)b * in_batch_stride + (long long)t * c_pitch + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (c < C && t < T) out[((long long)b * C + c) * T + t] = tile[tx][i];
  }
}

int cl_to_cf(const void* in, int in_type, float* out, int nb, int C, int T, long long in_batch_stride, int c_pitch, cudaStream_t s,
             const long long* lens) {
  dim3 grid(ceil_div(T, 32), ceil_div(C, 32), nb), block(32, 8);
  ProfScope ps(KC_MISC, s);
  if (in_type == 0) cl_to_cf_kernel<float><<<grid, block, 0, s>>>((const float*)in, out, C, T, in_batch_stride, c_pitch, lens);
  else cl_to_cf_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)in, out, C, T, in_batch_stride, c_pitch, lens);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
// log-mel pre/post passes (reference feature_extractor.py:207-214 zero-pad to 480000, torch.stft
// center/reflect framing, :104-109 per-item max-8 clamp and (x+4)/4)
// ================================================================================================
constexpr int kMelSamples = 480000, kMelPad = 200, kMelFrames = 3000, kMelBins = 80;

__global__ void mel_pad_kernel(const float* __restrict__ wav, long long wav_stride, int wav_cols,
                               const long long* __restrict__ lens, float* __restrict__ padded,
                               long long* __restrict__ mel_lens, float* __restrict__ item_max) {
  const int b = blockIdx.y;
  long long len = lens[b];
  len = len < 0 ? 0 : (len > kMelSamples ? kMelSamples : len);
  if (len > wav_cols) len = wav_cols;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (mel_lens) mel_lens[b] = (len + 159) / 160;
    item_max[b] = -INFINITY;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kMelSamples + 2 * kMelPad) return;
  int sidx = i - kMelPad;
  if (sidx < 0) sidx = -sidx;
  else if (sidx >= kMelSamples) sidx = 2 * (kMelSamples - 1) - sidx;
  padded[(long long)b * (kMelSamples + 2 * kMelPad) + i] = (sidx < len) ? wav[(long long)b * wav_stride + sidx] : 0.f;
}

int mel_pad(const float* wav, long long wav_stride, int wav_cols, const long long* lens, int nb, float* padded,
            long long* mel_lens, float* item_max, cudaStream_t s) {
  dim3 grid(ceil_div(kMelSamples + 2 * kMelPad, 256), nb);
  ProfScope ps(KC_MISC, s);
  mel_pad_kernel<<<grid, 256, 0, s>>>(wav, wav_stride, wav_cols, lens, padded, mel_lens, item_max);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(448) mel_frames_split_kernel(const float* __restrict__ padded, bf16* __restrict__ frames) {
  // one block = 8 frames; a thread splits 8 consecutive samples of one frame into the three bf16 terms and writes each
  // plane with one 16-byte store (frame starts are 640 bytes apart, so the two float4 loads are aligned)
  const int f = threadIdx.x / 56, g = threadIdx.x - f * 56;
  const int t = blockIdx.x * 8 + f, b = blockIdx.y;
  if (t >= kMelFrames) return;
  float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (g < 50) {
    const float* src = padded + (long long)b * (kMelSamples + 2 * kMelPad) + (long long)t * 160 + g * 8;
    const float4 u = *reinterpret_cast<const float4*>(src), v = *reinterpret_cast<const float4*>(src + 4);
    x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
  }
  __align__(16) bf16 a1[8], a2[8], a3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a1[i] = __float2bfloat16_rn(x[i]);
    const float r1 = x[i] - __bfloat162float(a1[i]);
    a2[i] = __float2bfloat16_rn(r1);
    a3[i] = __float2bfloat16_rn(r1 - __bfloat162float(a2[i]));
  }
  bf16* o = frames + ((long long)b * kMelFrames + t) * (3 * 448) + g * 8;
  *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(a1);
  *reinterpret_cast<uint4*>(o + 448) = *reinterpret_cast<const uint4*>(a2);
  *reinterpret_cast<uint4*>(o + 896) = *reinterpret_cast<const uint4*>(a3);
}
int mel_frames_split(const float* padded, int nb, bf16* frames, cudaStream_t s) {
  SWC_REQUIRE(((uintptr_t)padded & 15) == 0 && ((uintptr_t)frames & 15) == 0, "mel_frames_split: buffers must be 16-byte aligned");
  dim3 grid(ceil_div(kMelFrames, 8), nb);
  ProfScope ps(KC_MISC, s);
  mel_frames_split_kernel<<<grid, 448, 0, s>>>(padded, frames);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename TO>
__global__ void mel_finalize_kernel(const float* __restrict__ logmel, const float* __restrict__ item_max,
                                    float* __restrict__ mel_cf, TO* __restrict__ mel_cl, int cl_pitch) {
  __shared__ float tile[32][kMelBins + 1];
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const float floor_v = item_max[b] - 8.0f;
  for (int i = threadIdx.x; i < 32 * kMelBins; i += blockDim.x) {
    const int r = i / kMelBins, m = i % kMelBins;
    const int t = t0 + r;
    float v = 0.f;
    if (t < kMelFrames) v = (fmaxf(logmel[((long long)b * kMelFrames + t) * kMelBins + m], floor_v) + 4.0f) / 4.0f;
    tile[r][m] = v;
  }
  __syncthreads();
  if (mel_cf) {
    for (int i = threadIdx.x; i < 32 * kMelBins; i += blockDim.x) {
      const int m = i / 32, r = i % 32;
      if (t0 + r < kMelFrames) mel_cf[((long long)b * kMelBins + m) * kMelFrames + t0 + r] = tile[r][m];
    }
  }
  if (mel_cl) {
    for (int i = threadIdx.x; i < 32 * cl_pitch; i += blockDim.x) {
      const int r = i / cl_pitch, m = i % cl_pitch;
      if (t0 + r < kMelFrames)
        mel_cl[((long long)b * kMelFrames + t0 + r) * cl_pitch + m] = from_f32<TO>(m < kMelBins ? tile[r][m] : 0.f);
    }
  }
}

int mel_finalize(const float* logmel, const float* item_max, int nb, float* mel_cf, void* mel_cl, int cl_type,
                 int cl_pitch, cudaStream_t s) {
  dim3 grid(ceil_div(kMelFrames, 32), nb);
  ProfScope ps(KC_MISC, s);
  if (cl_type == 0) mel_finalize_kernel<float><<<grid, 256, 0, s>>>(logmel, item_max, mel_cf, (float*)mel_cl, cl_pitch);
  else mel_finalize_kernel<bf16><<<grid, 256, 0, s>>>(logmel, item_max, mel_cf, (bf16*)mel_cl, cl_pitch);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
// iSTFT overlap-add with "same" padding (reference modules.py:861-884): n_fft 640, hop 160.
// out[s] = sum_t frames[t][p-160t] / sum_t w^2[p-160t],  p = s + 240, t in the <=4 overlapping frames.
// ================================================================================================
template <typename TAB>
__global__ void __launch_bounds__(256) istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ win_sq, int T_dense,
                                                        float* __restrict__ wav, long long wav_stride, const __grid_constant__ TAB tab) {
  // one thread = 4 consecutive output samples: hop, trim and frame length are multiples of 4, so the four samples sit in
  // the same (<= 4) frames and every access is a float4
  __shared__ __align__(16) float w2[640];
  for (int i = threadIdx.x; i < 640; i += blockDim.x) w2[i] = win_sq[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int sidx = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int T;
  long long row0;
  if constexpr (std::is_same<TAB, RaggedTable>::value) { T = tab.len[b]; row0 = tab.off[b]; }
  else { T = T_dense; row0 = (long long)b * T; }
  const int L = 160 * T;
  if (sidx >= L) return;
  const int p = sidx + 240;
  const int t_hi = min(T - 1, p / 160);
  const int t_lo = max(0, (p - 639 + 159) / 160);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), env = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = t_lo; t <= t_hi; ++t) {
    const int n = p - 160 * t;
    const float4 f = *reinterpret_cast<const float4*>(frames + (row0 + t) * 640 + n);
    const float4 w = *reinterpret_cast<const float4*>(w2 + n);
    acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
    env.x += w.x; env.y += w.y; env.z += w.z; env.w += w.w;
  }
  *reinterpret_cast<float4*>(wav + (long long)b * wav_stride + sidx) = make_float4(acc.x / env.x, acc.y / env.y, acc.z / env.z, acc.w / env.w);
}
int istft_ola(const float* frames, const float* win_sq, int nb, int T, float* wav, long long wav_stride, cudaStream_t s) {
  SWC_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)wav & 15) == 0 && wav_stride % 4 == 0, "istft_ola: buffers must be 16-byte aligned");
  dim3 grid(ceil_div(40 * T, 256), nb);
  ProfScope ps(KC_MISC, s);
  istft_ola_kernel<DenseRows><<<grid, 256, 0, s>>>(frames, win_sq, T, wav, wav_stride, DenseRows{});
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// packed frames: item b = rows [off[b], off[b] + len[b]) of `frames`, written to wav + b * wav_stride (160 len[b] samples)
int istft_ola_ragged(const float* frames, const float* win_sq, const RaggedTable& tab, float* wav, long long wav_stride, cudaStream_t s) {
  SWC_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)wav & 15) == 0 && wav_stride % 4 == 0, "istft_ola_ragged: buffers must be 16-byte aligned");
  SWC_REQUIRE(tab.nb > 0 && tab.nb <= kMaxRagged && tab.t_max > 0, "istft_ola_ragged: bad table");
  dim3 grid(ceil_div(40 * tab.t_max, 256), tab.nb);
  ProfScope ps(KC_MISC, s);
  istft_ola_kernel<RaggedTable><<<grid, 256, 0, s>>>(frames, win_sq, 0, wav, wav_stride, tab);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
// bf16x3 mode: split an fp32 GEMM operand into two bf16 planes (x = hi + lo up to 2^-17 relative)
// ================================================================================================
__global__ void __launch_bounds__(256) split_bf16_planes_kernel(const float* __restrict__ in, long long row_stride, long long batch_stride,
                                                                int rows, int cols8, bf16* __restrict__ planes, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % cols8);
  const long long br = i / cols8;
  const int r = (int)(br % rows);
  const long long b = br / rows;
  const float* src = in + b * batch_stride + (long long)r * row_stride + c8 * 8;
  const float4 u = *reinterpret_cast<const float4*>(src), v = *reinterpret_cast<const float4*>(src + 4);
  const float x[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
  __align__(16) bf16 hi[8], lo[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    hi[k] = __float2bfloat16_rn(x[k]);
    lo[k] = __float2bfloat16_rn(x[k] - __bfloat162float(hi[k]));
  }
  bf16* o = planes + br * (2ll * cols8 * 8) + c8 * 8;
  *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(o + cols8 * 8) = *reinterpret_cast<const uint4*>(lo);
}
int split_bf16_planes(const float* in, long long row_stride, long long batch_stride, int nb, int rows, int cols, bf16* planes,
                      cudaStream_t s) {
  SWC_REQUIRE(cols % 8 == 0 && row_stride % 4 == 0 && batch_stride % 4 == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)planes & 15) == 0,
              "split_bf16_planes: cols %d / strides must be multiples of 8 / 4 elements and the buffers 16-byte aligned", cols);
  const long long total = (long long)nb * rows * (cols / 8);
  if (total == 0) return 0;
  ProfScope ps(KC_MISC, s);
  split_bf16_planes_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, s>>>(in, row_stride, batch_stride, rows, cols / 8, planes, total);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// planes (rows, hi cols | lo cols) bf16 -> fp32 rows (hi + lo): the inverse view, used by the test entry points
__global__ void __launch_bounds__(256) merge_bf16_planes_kernel(const bf16* __restrict__ planes, float* __restrict__ out, int cols, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  out[i] = __bfloat162float(planes[r * 2 * cols + c]) + __bfloat162float(planes[r * 2 * cols + cols + c]);
}
int merge_bf16_planes(const bf16* planes, long long rows, int cols, float* out, cudaStream_t s) {
  const long long total = rows * cols;
  if (total == 0) return 0;
  ProfScope ps(KC_MISC, s);
  merge_bf16_planes_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, s>>>(planes, out, cols, total);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
// ragged pack / unpack of token rows (float4 / 8-byte granules; C is a multiple of 8)
// ================================================================================================
__global__ void pack_rows_kernel(const float* __restrict__ padded, float* __restrict__ packed, const __grid_constant__ RaggedTable tab,
                                 int t_pad, int C4) {
  const int b = blockIdx.y;
  const int n = tab.len[b];
  const float4* src = reinterpret_cast<const float4*>(padded) + (long long)b * t_pad * C4;
  float4* dst = reinterpret_cast<float4*>(packed) + (long long)tab.off[b] * C4;
  const long long total = (long long)n * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
template <typename T>
__global__ void unpack_rows_kernel(const T* __restrict__ packed, T* __restrict__ padded, const __grid_constant__ RaggedTable tab,
                                   int t_pad, int C8) {
  // one uint4 = 8 bf16 or 4 fp32; C8 = uint4 per row
  const int b = blockIdx.y;
  const long long valid = (long long)tab.len[b] * C8, total = (long long)t_pad * C8;
  const uint4* src = reinterpret_cast<const uint4*>(packed) + (long long)tab.off[b] * C8;
  uint4* dst = reinterpret_cast<uint4*>(padded) + (long long)b * t_pad * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = i < valid ? src[i] : make_uint4(0, 0, 0, 0);
}
// the same for rows of `row_bytes` (a multiple of 16) of any type, and the rows between the items' ends and the next item's
// start (off[b] + len[b] .. off[b + 1]) are written as zero: the packed Vocos input, whose 7-tap embedding convolution reads
// up to three rows across an item's edges
__global__ void pack_rows_gap_kernel(const uint4* __restrict__ padded, uint4* __restrict__ packed, const __grid_constant__ RaggedTable tab,
                                     int t_pad, int G) {
  const int b = blockIdx.y;
  const long long valid = (long long)tab.len[b] * G, total = (long long)(tab.off[b + 1] - tab.off[b]) * G;
  const uint4* src = padded + (long long)b * t_pad * G;
  uint4* dst = packed + (long long)tab.off[b] * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = i < valid ? src[i] : make_uint4(0, 0, 0, 0);
}
int pack_rows_gap(const void* padded, void* packed, const RaggedTable& tab, int t_pad, int row_bytes, cudaStream_t s) {
  SWC_REQUIRE(row_bytes % 16 == 0 && tab.nb > 0 && tab.t_max <= t_pad, "pack_rows_gap: bad shape");
  const int G = row_bytes / 16;
  dim3 grid((unsigned)std::max(1, std::min(64, ceil_div((tab.t_max + 8) * G, 256))), tab.nb);
  ProfScope ps(KC_MISC, s);
  pack_rows_gap_kernel<<<grid, 256, 0, s>>>((const uint4*)padded, (uint4*)packed, tab, t_pad, G);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int pack_rows(const float* padded, float* packed, const RaggedTable& tab, int t_pad, int C, cudaStream_t s) {
  SWC_REQUIRE(C % 4 == 0 && tab.nb > 0, "pack_rows: bad shape");
  dim3 grid((unsigned)std::max(1, std::min(64, ceil_div(tab.t_max * (C / 4), 256))), tab.nb);
  ProfScope ps(KC_MISC, s);
  pack_rows_kernel<<<grid, 256, 0, s>>>(padded, packed, tab, t_pad, C / 4);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int unpack_rows(const void* packed, void* padded, int dtype, const RaggedTable& tab, int t_pad, int C, cudaStream_t s) {
  SWC_REQUIRE(C % 8 == 0 && tab.nb > 0, "unpack_rows: bad shape");
  const int c8 = dtype == 0 ? C / 4 : C / 8;
  dim3 grid((unsigned)std::max(1, std::min(64, ceil_div(t_pad * c8, 256))), tab.nb);
  ProfScope ps(KC_MISC, s);
  if (dtype == 0) unpack_rows_kernel<float><<<grid, 256, 0, s>>>((const float*)packed, (float*)padded, tab, t_pad, c8);
  else unpack_rows_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)packed, (bf16*)padded, tab, t_pad, c8);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
__global__ void fill_f32_kernel(float* p, float v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_f32(float* p, float v, long long n, cudaStream_t s) {
  ProfScope ps(KC_MISC, s);
  fill_f32_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, s>>>(p, v, n);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
__global__ void cvt_bf16_kernel(const float* in, bf16* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
int convert_f32_to_bf16(const float* in, bf16* out, long long n, cudaStream_t s) {
  ProfScope ps(KC_MISC, s);
  cvt_bf16_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, s>>>(in, out, n);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace swc

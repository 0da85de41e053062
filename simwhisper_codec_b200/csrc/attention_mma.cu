// bf16 flash attention on the legacy tensor-core path (mma.sync.m16n8k16, fp32 accumulate) — the
// first tensor-core version of the fused MHA (reference audiocodec/nn/modules.py:145-187): non-causal,
// head_dim 64, keys >= lens[b] masked, q pre-scaled.  One CTA = 128 queries of one (batch, head), 8 warps
// x 16 query rows; K/V stream through a double-buffered, XOR-swizzled shared-memory ring filled by
// cp.async; scores, the online softmax and the P fragments never leave registers.
#include "kernels.cuh"

namespace swc {

namespace {

constexpr int QT = 128, KT = 64, HD = 64;
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tile rows are 64 bf16 = 128 B = 8 chunks of 16 B; chunk c of row r lives at chunk (c ^ (r & 7))
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 2) attention_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                    const long long* __restrict__ lens, int T, int H) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;                        // 128 x 128 B
  uint8_t* sK = smem + QT * 128;             // 2 x 64 x 128 B
  uint8_t* sV = sK + 2 * KT * 128;           // 2 x 64 x 128 B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int D3 = 3 * H * HD, DO = H * HD;
  long long len_ll = lens ? lens[b] : T;
  const int len = (int)(len_ll > T ? T : (len_ll < 0 ? 0 : len_ll));
  const bf16* base = qkv + (long long)b * T * D3;
  bf16* obase = out + (long long)b * T * DO + h * HD;

  if (q0 >= len) {   // tile of padded queries: defined zero output
    for (int i = tid; i < QT * 8; i += kThreads) {
      const int r = i >> 3, c = i & 7;
      if (q0 + r < T) *reinterpret_cast<uint4*>(obase + (long long)(q0 + r) * DO + c * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  // ---- async loads: Q (once) and K/V tile 0
  for (int i = tid; i < QT * 8; i += kThreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = q0 + r < T;
    cp_async16(smem_u32(sQ) + swz(r, c), base + (long long)(ok ? q0 + r : 0) * D3 + h * HD + c * 8, ok);
  }
  auto load_kv = [&](int kt, int buf) {
    const int k0 = kt * KT;
    for (int i = tid; i < KT * 8; i += kThreads) {
      const int r = i >> 3, c = i & 7;
      const bool ok = k0 + r < len;           // masked keys are zero-filled (no NaN can enter P*V)
      const bf16* src = base + (long long)(ok ? k0 + r : 0) * D3 + c * 8;
      cp_async16(smem_u32(sK) + buf * KT * 128 + swz(r, c), src + (H + h) * HD, ok);
      cp_async16(smem_u32(sV) + buf * KT * 128 + swz(r, c), src + (2 * H + h) * HD, ok);
    }
  };
  load_kv(0, 0);
  cp_commit();

  const int n_kt = (len + KT - 1) / KT;
  const int g = lane >> 2, tq = lane & 3;
  const int wrow = warp * 16;

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) {
      load_kv(kt + 1, buf ^ 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = wrow + (lane & 15), c = ks * 2 + (lane >> 4);
        ldsm_x4(smem_u32(sQ) + swz(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t kb = smem_u32(sK) + buf * KT * 128, vb = smem_u32(sV) + buf * KT * 128;

    // ---- S = Q K^T : 16 x 64 per warp
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {     // pairs of 8-key tiles
        uint32_t b0, b1, b2, b3;
        const int r = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int c = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kb + swz(r, c), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qf[ks], b0, b1);
        mma_bf16(s[2 * np + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask (only the last tile can hold masked keys) + online softmax
    const int k0 = kt * KT;
    if (k0 + KT > len) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = k0 + nt * 8 + tq * 2;
        if (key >= len) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (key + 1 >= len) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);     // finite: key k0 < len is valid in every tile
    const float c0 = exp2f((m0 - mn0) * kLog2e), c1 = exp2f((m1 - mn1) * kLog2e);
    m0 = mn0; m1 = mn1;
    const float ms0 = mn0 * kLog2e, ms1 = mn1 * kLog2e;
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(fmaf(s[nt][0], kLog2e, -ms0)), p1 = exp2f(fmaf(s[nt][1], kLog2e, -ms0));
      const float p2 = exp2f(fmaf(s[nt][2], kLog2e, -ms1)), p3 = exp2f(fmaf(s[nt][3], kLog2e, -ms1));
      ps0 += p0 + p1;
      ps1 += p2 + p3;
      // C fragment of key tile nt -> A fragment of k-step nt/2 (keys 16*(nt/2) + ...)
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l0 = l0 * c0 + ps0;
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
    // ---- O += P V : k = 64 keys (4 steps), n = 64 dims (8 tiles)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {     // pairs of 8-dim tiles
        uint32_t b0, b1, b2, b3;
        const int r = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int c = dp * 2 + (lane >> 4);
        ldsm_x4_t(vb + swz(r, c), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pf[ks], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();   // everyone is done with buffer `buf` before it is refilled two iterations later
  }

  // row sums live spread over the 4 lanes of a quad
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int r0 = q0 + wrow + g, r1 = r0 + 8;
  const float i0 = (r0 < len && l0 > 0.f) ? 1.0f / l0 : 0.f;   // padded query rows -> 0
  const float i1 = (r1 < len && l1 > 0.f) ? 1.0f / l1 : 0.f;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int d = dt * 8 + tq * 2;
    if (r0 < T) *reinterpret_cast<uint32_t*>(obase + (long long)r0 * DO + d) = pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(obase + (long long)r1 * DO + d) = pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------
// bf16x3 mode: the same flash attention at fp32-class accuracy.  q, k, v arrive as two bf16 planes each (x = hi + lo up to
// 2^-17, split_bf16_planes of the fp32 qkv rows: plane row = [hi (3 H 64) | lo (3 H 64)]); every product is evaluated as
// hi*hi + hi*lo + lo*hi with fp32 accumulation: S = Qh Kh^T + Ql Kh^T + Qh Kl^T, and the fp32 probabilities are split the
// same way in registers for O += Ph Vh + Pl Vh + Ph Vl.  Softmax, row sums and the output stay fp32.
// ------------------------------------------------------------------------------------------------
template <bool RAGGED>
__global__ void __launch_bounds__(kThreads, 2) attention_mma_x3_kernel(const bf16* __restrict__ planes, float* __restrict__ out,
                                                                       const long long* __restrict__ lens, int T, int H,
                                                                       const __grid_constant__ RaggedTable tab) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;                        // [hi, lo] x 128 x 128 B
  uint8_t* sK = smem + 2 * QT * 128;         // [2 buffers][hi, lo] x 64 x 128 B
  uint8_t* sV = sK + 4 * KT * 128;           // [2 buffers][hi, lo] x 64 x 128 B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int D3 = 3 * H * HD, DO = H * HD, PR = 2 * D3;      // PR: plane row length
  // RAGGED: rows are packed (item b = rows [off[b], off[b] + len[b])); rows >= len do not exist and nothing is written for them
  int len;
  long long row0;
  if constexpr (RAGGED) {
    len = tab.len[b];
    row0 = tab.off[b];
  } else {
    long long len_ll = lens ? lens[b] : T;
    len = (int)(len_ll > T ? T : (len_ll < 0 ? 0 : len_ll));
    row0 = (long long)b * T;
  }
  const int Tq = RAGGED ? len : T;           // query rows that exist
  const bf16* base = planes + row0 * PR;
  float* obase = out + row0 * DO + h * HD;

  if (q0 >= len) {   // tile of padded queries: defined zero output
    if constexpr (RAGGED) return;
    for (int i = tid; i < QT * 16; i += kThreads) {
      const int r = i >> 4, c = i & 15;
      if (q0 + r < T) *reinterpret_cast<float4*>(obase + (long long)(q0 + r) * DO + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }

  for (int i = tid; i < 2 * QT * 8; i += kThreads) {
    const int pl = i / (QT * 8), j = i - pl * (QT * 8), r = j >> 3, c = j & 7;
    const bool ok = q0 + r < Tq;
    cp_async16(smem_u32(sQ) + pl * QT * 128 + swz(r, c), base + (long long)(ok ? q0 + r : 0) * PR + pl * D3 + h * HD + c * 8, ok);
  }
  auto load_kv = [&](int kt, int buf) {
    const int k0 = kt * KT;
    for (int i = tid; i < 2 * KT * 8; i += kThreads) {
      const int pl = i / (KT * 8), j = i - pl * (KT * 8), r = j >> 3, c = j & 7;
      const bool ok = k0 + r < len;           // masked keys are zero-filled (no NaN can enter P*V)
      const bf16* src = base + (long long)(ok ? k0 + r : 0) * PR + pl * D3 + c * 8;
      cp_async16(smem_u32(sK) + (buf * 2 + pl) * KT * 128 + swz(r, c), src + (H + h) * HD, ok);
      cp_async16(smem_u32(sV) + (buf * 2 + pl) * KT * 128 + swz(r, c), src + (2 * H + h) * HD, ok);
    }
  };
  load_kv(0, 0);
  cp_commit();

  const int n_kt = (len + KT - 1) / KT;
  const int g = lane >> 2, tq = lane & 3;
  const int wrow = warp * 16;

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) {
      load_kv(kt + 1, buf ^ 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    const uint32_t kh = smem_u32(sK) + (buf * 2) * KT * 128, kl = kh + KT * 128;
    const uint32_t vh = smem_u32(sV) + (buf * 2) * KT * 128, vl = vh + KT * 128;

    // ---- S = Qh Kh^T + Ql Kh^T + Qh Kl^T : 16 x 64 per warp (the Q fragments are re-read per k-step: registers)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t qh[4], ql[4];
      {
        const int r = wrow + (lane & 15), c = ks * 2 + (lane >> 4);
        ldsm_x4(smem_u32(sQ) + swz(r, c), qh[0], qh[1], qh[2], qh[3]);
        ldsm_x4(smem_u32(sQ) + QT * 128 + swz(r, c), ql[0], ql[1], ql[2], ql[3]);
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {     // pairs of 8-key tiles
        uint32_t b0, b1, b2, b3;
        const int r = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int c = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kl + swz(r, c), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qh, b0, b1);
        mma_bf16(s[2 * np + 1], qh, b2, b3);
        ldsm_x4(kh + swz(r, c), b0, b1, b2, b3);
        mma_bf16(s[2 * np], ql, b0, b1);
        mma_bf16(s[2 * np + 1], ql, b2, b3);
        mma_bf16(s[2 * np], qh, b0, b1);
        mma_bf16(s[2 * np + 1], qh, b2, b3);
      }
    }
    const int k0 = kt * KT;
    if (k0 + KT > len) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = k0 + nt * 8 + tq * 2;
        if (key >= len) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (key + 1 >= len) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);     // finite: key k0 < len is valid in every tile
    const float c0 = exp2f((m0 - mn0) * kLog2e), c1 = exp2f((m1 - mn1) * kLog2e);
    m0 = mn0; m1 = mn1;
    const float ms0 = mn0 * kLog2e, ms1 = mn1 * kLog2e;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {      // probabilities stay fp32 in the score registers
      s[nt][0] = exp2f(fmaf(s[nt][0], kLog2e, -ms0)); s[nt][1] = exp2f(fmaf(s[nt][1], kLog2e, -ms0));
      s[nt][2] = exp2f(fmaf(s[nt][2], kLog2e, -ms1)); s[nt][3] = exp2f(fmaf(s[nt][3], kLog2e, -ms1));
      ps0 += s[nt][0] + s[nt][1];
      ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + ps0;
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
    // ---- O += Ph Vh + Pl Vh + Ph Vl : k = 64 keys (4 steps), n = 64 dims (8 tiles)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {        // C fragments of key tiles 2ks, 2ks+1 -> A fragment of k-step ks
        const float a = s[2 * ks + (e >> 1)][(e & 1) * 2], bb = s[2 * ks + (e >> 1)][(e & 1) * 2 + 1];
        const __nv_bfloat162 hi = __floats2bfloat162_rn(a, bb);
        const float2 hf = __bfloat1622float2(hi);
        ph[e] = *reinterpret_cast<const uint32_t*>(&hi);
        pl[e] = pack_bf16(a - hf.x, bb - hf.y);
      }
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {     // pairs of 8-dim tiles
        uint32_t b0, b1, b2, b3;
        const int r = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int c = dp * 2 + (lane >> 4);
        ldsm_x4_t(vl + swz(r, c), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], ph, b0, b1);
        mma_bf16(o[2 * dp + 1], ph, b2, b3);
        ldsm_x4_t(vh + swz(r, c), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pl, b0, b1);
        mma_bf16(o[2 * dp + 1], pl, b2, b3);
        mma_bf16(o[2 * dp], ph, b0, b1);
        mma_bf16(o[2 * dp + 1], ph, b2, b3);
      }
    }
    __syncthreads();   // everyone is done with buffer `buf` before it is refilled two iterations later
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int r0 = q0 + wrow + g, r1 = r0 + 8;
  const float i0 = (r0 < len && l0 > 0.f) ? 1.0f / l0 : 0.f;   // padded query rows -> 0
  const float i1 = (r1 < len && l1 > 0.f) ? 1.0f / l1 : 0.f;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int d = dt * 8 + tq * 2;
    if (r0 < Tq) *reinterpret_cast<float2*>(obase + (long long)r0 * DO + d) = make_float2(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < Tq) *reinterpret_cast<float2*>(obase + (long long)r1 * DO + d) = make_float2(o[dt][2] * i1, o[dt][3] * i1);
  }
}

constexpr int kSmemBytesX3 = 2 * QT * 128 + 8 * KT * 128;
constexpr int kSmemBytes = QT * 128 + 4 * KT * 128;

}  // namespace

int attention_mma(const bf16* qkv, bf16* out, const long long* lens, int nb, int T, int H, cudaStream_t s) {
  dim3 grid(ceil_div(T, QT), H, nb);
  SWC_TRY(ensure_dynamic_smem((const void*)attention_mma_kernel, kSmemBytes));
  ProfScope ps(KC_ATTN, s);
  attention_mma_kernel<<<grid, kThreads, kSmemBytes, s>>>(qkv, out, lens, T, H);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attention_mma_x3(const bf16* planes, float* out, const long long* lens, int nb, int T, int H, cudaStream_t s) {
  dim3 grid(ceil_div(T, QT), H, nb);
  SWC_TRY(ensure_dynamic_smem((const void*)attention_mma_x3_kernel<false>, kSmemBytesX3));
  ProfScope ps(KC_ATTN, s);
  attention_mma_x3_kernel<false><<<grid, kThreads, kSmemBytesX3, s>>>(planes, out, lens, T, H, RaggedTable{});
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attention_mma_x3_ragged(const bf16* planes, float* out, const RaggedTable& tab, int H, cudaStream_t s) {
  SWC_REQUIRE(tab.nb > 0 && tab.nb <= kMaxRagged && tab.total > 0 && tab.t_max > 0, "attention_mma_x3_ragged: bad table");
  dim3 grid(ceil_div(tab.t_max, QT), H, tab.nb);
  SWC_TRY(ensure_dynamic_smem((const void*)attention_mma_x3_kernel<true>, kSmemBytesX3));
  ProfScope ps(KC_ATTN, s);
  attention_mma_x3_kernel<true><<<grid, kThreads, kSmemBytesX3, s>>>(planes, out, nullptr, tab.t_max, H, tab);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace swc

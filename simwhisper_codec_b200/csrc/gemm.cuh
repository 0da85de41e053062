// Problem descriptor shared by the two GEMM back ends (fp32 SIMT and bf16 tcgen05) and the fused
// epilogues.  Every dense contraction of the codec is expressed as
//
//     D[b, m, n] = sum_{tap} sum_{k < tap_k} A[b, m + tap_row[tap], tap_col[tap] + k] * W[n, tap*tap_k + k]
//
// with A a channel-last activation (rows outside [0, a_rows) read as zero) — i.e. an implicit-GEMM
// 1-D convolution whose im2col is only a table of (row shift, column offset) per tap.  Plain linear
// layers are the 1-tap case.
#pragma once
#include "common.cuh"

namespace swc {

enum EpiKind { EPI_STORE = 0, EPI_POWER = 1, EPI_LOGMEL = 2, EPI_FSQ = 3, EPI_HEAD = 4 };

constexpr int kMaxTaps = 24;     // 7-tap convolutions x 3 split-bf16 products

struct FsqConst {        // per-dimension constants of one 4-dim FSQ group (reference quantizer.py:129-179)
  float scale[4];        // (L-1)/2 * (1-eps)
  float offset[4];       // 0.5 for even L else 0
  float shift[4];        // tan(offset/scale)
  float half[4];         // L // 2
  int base[4];           // mixed-radix base
  int levels[4];
};

struct EpiParams {
  // EPI_STORE: out = (act(acc + bias)) * gamma + residual
  const float* bias;        // [N] or null
  const float* gamma;       // [N] or null
  const float* residual;    // fp32, may alias out
  long long res_row_stride, res_batch_stride;
  int act;                  // 0 none, 1 exact-erf GELU (erff), 2 erf GELU via gelu_fast (bf16 outputs)
  void* out;                // TO
  long long out_row_stride, out_batch_stride;
  int out_row_mul, out_row_off;   // output row = m * mul + off (deconv parity interleave)
  bf16* out2;               // optional second bf16 copy with the same addressing (bf16 mode)
  int out_planes;           // bf16x3 mode: write the result as two bf16 planes (hi at column n, lo at column N + n) of a bf16
                            // output with 2N-wide rows: the split operand of the next three-product GEMM (pair kernel only)
  // EPI_LOGMEL
  float* item_max;          // [nb], initialised to -inf
  // EPI_FSQ (N == 32)
  const long long* lens;    // [nb] valid code frames
  int nb;                   // batches (codes are laid out (G, nb, m_rows))
  int* codes;               // (G, nb, m_rows) int32 or null
  float* zq_cf;             // (nb, 32, m_rows) fp32 channels-first or null
  float* latent_cf;         // (nb, 32, m_rows) fp32 channels-first or null (pre-quantisation)
  float* zq_cl;             // (nb, m_rows, 32) fp32 channel-last or null (feeds the up-sampler)
  FsqConst fsq;
};

struct GemmDesc {
  const void* A;            // TA
  long long a_row_stride;   // elements
  long long a_batch_stride; // elements
  int a_rows;               // valid rows per batch
  int a_cols;               // logical row width (for the tensor map)
  int a_planes;             // bf16x3 mode: A already holds (hi | lo) bf16 planes of a_cols columns each (strides in bf16 elements)
  int m_rows;               // output rows per batch
  int nb;                   // batches
  int n_taps;
  int tap_row[kMaxTaps];
  int tap_col[kMaxTaps];
  int tap_k;                // multiple of 16 (SIMT) / 64 (tcgen05)
  const void* W;            // [N_pad, K] row-major, K = n_taps * tap_k
  const void* W3;           // bf16x3 mode: [N_pad, 3K] bf16 planes (hi | lo | hi), else null
  int N;                    // logical output columns
  int w_rows;               // rows present in W (>= N)
  EpiParams epi;
};

// ------------------------------------------------------------------------------------------------
// FSQ on 4 consecutive latent channels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int fsq_quantize4(const FsqConst& c, const float* x, float* dq) {
  int idx = 0;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    float comp = c.scale[d] * tanhf(x[d] + c.shift[d]) - c.offset[d];
    float r = rintf(comp);                       // ties-to-even like torch.round
    dq[d] = r / c.half[d];
    idx += (int)(r + c.half[d]) * c.base[d];
  }
  return idx;
}

// ------------------------------------------------------------------------------------------------
// epilogue on 8 consecutive columns [n0, n0+8) of row m of batch b
// ------------------------------------------------------------------------------------------------
// exact-erf GELU to ~2.5e-5 absolute: erf(x/sqrt2) ~ tanh(x (c0 + c1 x^2 + c2 x^4)), one MUFU.TANH.
// Used only where the result is rounded to bf16 (half-ulp there is >= 1e-3 relative).
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 36.0f);
  const float poly = fmaf(fmaf(-0.000351517274f, x2, 0.0370056493f), x2, 0.79750788f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * poly));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// two GELUs per instruction on the packed-fp32 pipe (fma.rn.f32x2 / mul.rn.f32x2, sm_100), for results that are rounded
// to bf16: erf(x/sqrt2) ~ tanh(x (a + b x^2)) with the minimax pair a, b (max abs error of the GELU 2.7e-4, a fifteenth of
// the bf16 half-ulp at 1).  The cubic is monotone, so no clamp of x^2 is needed (tanh.approx saturates), and everything
// but the two MUFU.TANH is packed: 5 + 2 issue slots per pair instead of 12 for the three-term form of gelu_fast.
__device__ __forceinline__ unsigned long long gelu_fast2(unsigned long long x) {
  unsigned long long x2, u, ca, cb, half, t, hx, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(cb) : "f"(0.03470089f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ca) : "f"(0.80015708f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(half) : "f"(0.5f));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(x2) : "l"(x));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(u) : "l"(cb), "l"(x2), "l"(ca));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(x), "l"(u));
  float ta, tb;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(ta), "=f"(tb) : "l"(u));
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ta));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(tb));
  asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(ta), "f"(tb));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hx) : "l"(x), "l"(half));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(hx), "l"(t), "l"(hx));
  return r;
}
__device__ __forceinline__ void gelu_fast2(float& a, float& b) {
  unsigned long long x;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
  x = gelu_fast2(x);
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x));
}

template <int KIND, typename TO>
__device__ __forceinline__ void epi_apply(const EpiParams& p, int b, int m, int n0, int N, int m_rows,
                                          float (&v)[8], const float* bias8, const float* gamma8) {
  // bias8 / gamma8 point at the 8 per-column values of this chunk (global or shared memory) or are null
  if constexpr (KIND == EPI_STORE) {
    const bool full = (n0 + 8 <= N);
    if (bias8) {
      if (full) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias8), b1 = *reinterpret_cast<const float4*>(bias8 + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      } else {
        for (int j = 0; j < 8; ++j) if (n0 + j < N) v[j] += bias8[j];
      }
    }
    if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
    } else if (p.act == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gelu_fast(v[j]);
    }
    if (gamma8) {
      if (full) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma8), g1 = *reinterpret_cast<const float4*>(gamma8 + 4);
        v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w; v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
      } else {
        for (int j = 0; j < 8; ++j) if (n0 + j < N) v[j] *= gamma8[j];
      }
    }
    const long long orow = (long long)m * p.out_row_mul + p.out_row_off;
    if (p.residual) {
      const float* r = p.residual + (long long)b * p.res_batch_stride + orow * p.res_row_stride + n0;
      if (full) {
        float rv[8];
        load8(r, rv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += rv[j];
      } else {
        for (int j = 0; j < 8; ++j) if (n0 + j < N) v[j] += r[j];
      }
    }
    const long long off = (long long)b * p.out_batch_stride + orow * p.out_row_stride + n0;
    TO* o = reinterpret_cast<TO*>(p.out) + off;
    if (full) {
      store8(o, v);
      if (p.out2) store8(p.out2 + off, v);
    } else {
      for (int j = 0; j < 8; ++j) if (n0 + j < N) {
        o[j] = from_f32<TO>(v[j]);
        if (p.out2) p.out2[off + j] = __float2bfloat16_rn(v[j]);
      }
    }
  } else if constexpr (KIND == EPI_POWER) {
    // columns are interleaved (Re_k, Im_k); write |X_k|^2 to column k (fp32)
    float* o = reinterpret_cast<float*>(p.out) + (long long)b * p.out_batch_stride + (long long)m * p.out_row_stride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + 2 * j;
      if (n + 1 < N) o[n >> 1] = v[2 * j] * v[2 * j] + v[2 * j + 1] * v[2 * j + 1];
    }
  } else if constexpr (KIND == EPI_LOGMEL) {
    float* o = reinterpret_cast<float*>(p.out) + (long long)b * p.out_batch_stride + (long long)m * p.out_row_stride + n0;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (n0 + j < N) {
        float l = log10f(fmaxf(v[j], 1e-10f));
        o[j] = l;
        mx = fmaxf(mx, l);
      }
    }
    if (mx > -INFINITY) atomic_max_float(p.item_max + b, mx);
  } else if constexpr (KIND == EPI_FSQ) {
    // two FSQ groups per 8 columns
    const bool valid = (long long)m < p.lens[b];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = (n0 >> 2) + h;
      float x[4], dq[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) x[d] = v[4 * h + d] + (bias8 ? bias8[4 * h + d] : 0.0f);
      int idx = fsq_quantize4(p.fsq, x, dq);
      if (!valid) { idx = 0; dq[0] = dq[1] = dq[2] = dq[3] = 0.0f; }
      if (p.codes) p.codes[((long long)g * p.nb + b) * m_rows + m] = idx;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int c = n0 + 4 * h + d;
        const long long cf = ((long long)b * 32 + c) * m_rows + m;
        if (p.zq_cf) p.zq_cf[cf] = dq[d];
        if (p.latent_cf) p.latent_cf[cf] = x[d];
        if (p.zq_cl) p.zq_cl[((long long)b * m_rows + m) * 32 + c] = dq[d];
      }
    }
  } else if constexpr (KIND == EPI_HEAD) {
    // columns interleaved (log-magnitude_k, phase_k) -> S_k = min(exp(mag),100) * (cos p, sin p), fp32
    float* o = reinterpret_cast<float*>(p.out) + (long long)b * p.out_batch_stride + (long long)m * p.out_row_stride + n0;
    float r[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lm = v[2 * j], ph = v[2 * j + 1];
      if (bias8) { lm += bias8[2 * j]; ph += bias8[2 * j + 1]; }
      float mag = fminf(expf(lm), 100.0f);
      float s, c;
      sincosf(ph, &s, &c);
      r[2 * j] = mag * c;
      r[2 * j + 1] = mag * s;
    }
    if (p.out2) {   // tensor-core iDFT: S = s1 + s2 as two bf16 planes [row][s1 (N) | s2 (N)] in the same bytes as the fp32 row
      bf16* o2 = p.out2 + 2 * ((long long)b * p.out_batch_stride + (long long)m * p.out_row_stride) + n0;
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hi[j] = __bfloat162float(__float2bfloat16_rn(r[j]));
        lo[j] = r[j] - hi[j];
      }
      store8(o2, hi);
      store8(o2 + N, lo);
    } else {
      store8(o, r);   // the S buffer is padded to whole K slabs; pad columns meet zero iDFT rows
    }
  }
}

// launchers ----------------------------------------------------------------------------------------
// type codes: 0 = fp32, 1 = bf16
int gemm_simt(const GemmDesc& d, int kind, int a_type, int out_type, cudaStream_t s);
int gemm_tc(const GemmDesc& d, int kind, int out_type, int num_sms, cudaStream_t s);
// second-generation tcgen05 kernel (gemm_tc2.cu): CTA-pair MMA + TMA-store epilogue for plain EPI_STORE problems.
// variant 0 = first-generation kernel only, 1 = single-CTA tiles + TMA store, 2 = CTA pairs (default).
bool gemm_tc2_eligible(const GemmDesc& d);
int gemm_tc2(const GemmDesc& d, int out_type, int num_sms, int variant, cudaStream_t s);
void set_gemm_variant(int v);
int get_gemm_variant();

}  // namespace swc

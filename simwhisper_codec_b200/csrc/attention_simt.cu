// fp32 flash-style multi-head attention on CUDA cores (the fp32 parity mode; reference
// audiocodec/nn/modules.py:145-187).  Non-causal, head_dim 64, keys >= lens[b] masked out.  The
// reference's additive "+1 on valid pairs" (modules.py:142) is softmax-invariant and is dropped.
// One CTA = 64 queries of one (batch, head); K/V stream through shared memory in 64-key tiles with an
// online softmax, so the (T x T) score matrix is never materialised.
#include "kernels.cuh"

namespace swc {

namespace {
constexpr int QT = 64, KT = 64, HD = 64;
constexpr int PP = 68;   // padded pitch of the query and probability tiles (conflict-free row broadcast)
constexpr int kSmemBytes = (2 * QT * PP + HD * KT + KT * HD) * (int)sizeof(float);

template <typename T>
__global__ void __launch_bounds__(256) attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                             const long long* __restrict__ lens, int Tlen, int H) {
  extern __shared__ __align__(16) float smem[];
  float (*Qs)[PP] = reinterpret_cast<float(*)[PP]>(smem);                          // [query][d]
  float (*Kt)[KT] = reinterpret_cast<float(*)[KT]>(smem + QT * PP);                // [d][key]
  float (*Vs)[HD] = reinterpret_cast<float(*)[HD]>(smem + QT * PP + HD * KT);      // [key][d]
  float (*Ps)[PP] = reinterpret_cast<float(*)[PP]>(smem + QT * PP + HD * KT + KT * HD);  // [query][key]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * QT;
  const int h = blockIdx.y, b = blockIdx.z;
  const int D3 = 3 * H * HD;
  long long len = lens ? lens[b] : Tlen;
  if (len > Tlen) len = Tlen;
  const T* base = qkv + (long long)b * Tlen * D3;
  T* obase = out + (long long)b * Tlen * (H * HD) + h * HD;

  if (q0 >= len) {   // whole query tile is padding: defined (zero) output, no work
    for (int i = tid; i < QT * HD; i += 256) {
      const int r = i / HD, d = i % HD;
      if (q0 + r < Tlen) obase[(long long)(q0 + r) * (H * HD) + d] = from_f32<T>(0.f);
    }
    return;
  }

  // loader mapping: row = tid % 64, 16-wide d chunk = tid / 64
  const int lr = tid & 63, lc = (tid >> 6) * 16;
  {
    const int t = q0 + lr;
    float v[16];
    if (t < Tlen) {
      load8(base + (long long)t * D3 + h * HD + lc, *reinterpret_cast<float(*)[8]>(v));
      load8(base + (long long)t * D3 + h * HD + lc + 8, *reinterpret_cast<float(*)[8]>(v + 8));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) Qs[lr][lc + j] = v[j];
  }

  float o[4][4], mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY;
    lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  const int n_kt = (int)((len + KT - 1) / KT);
  for (int kt = 0; kt < n_kt; ++kt) {
    const int k0 = kt * KT;
    __syncthreads();   // previous tile fully consumed (also orders the Q stores before first use)
    {
      const int t = k0 + lr;
      float kv[16], vv[16];
      if (t < len) {
        const T* kp = base + (long long)t * D3 + (H + h) * HD + lc;
        const T* vp = base + (long long)t * D3 + (2 * H + h) * HD + lc;
        load8(kp, *reinterpret_cast<float(*)[8]>(kv));
        load8(kp + 8, *reinterpret_cast<float(*)[8]>(kv + 8));
        load8(vp, *reinterpret_cast<float(*)[8]>(vv));
        load8(vp + 8, *reinterpret_cast<float(*)[8]>(vv + 8));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { kv[j] = 0.f; vv[j] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { Kt[lc + j][lr] = kv[j]; Vs[lr][lc + j] = vv[j]; }
    }
    __syncthreads();

    // S = Q K^T for rows ty*4.., keys tx*4..
    float sacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      const float4 kk = *reinterpret_cast<const float4*>(&Kt[d][tx * 4]);
      float qq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qq[i] = Qs[ty * 4 + i][d];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sacc[i][0] = fmaf(qq[i], kk.x, sacc[i][0]);
        sacc[i][1] = fmaf(qq[i], kk.y, sacc[i][1]);
        sacc[i][2] = fmaf(qq[i], kk.z, sacc[i][2]);
        sacc[i][3] = fmaf(qq[i], kk.w, sacc[i][3]);
      }
    }
    // mask + online softmax (row statistics shared by the 16 lanes with equal ty)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + tx * 4 + j >= len) sacc[i][j] = -INFINITY;
        mx = fmaxf(mx, sacc[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(mrow[i], mx);          // finite: every tile has >= 1 valid key
      const float corr = expf(mrow[i] - mnew);
      float ps = 0.f;
      float pv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { pv[j] = expf(sacc[i][j] - mnew); ps += pv[j]; }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      lrow[i] = lrow[i] * corr + ps;
      mrow[i] = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
      *reinterpret_cast<float4*>(&Ps[ty * 4 + i][tx * 4]) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    }
    __syncthreads();
    // O += P V for rows ty*4.., dims tx*4..
#pragma unroll 8
    for (int c = 0; c < KT; ++c) {
      const float4 vv = *reinterpret_cast<const float4*>(&Vs[c][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = Ps[ty * 4 + i][c];
        o[i][0] = fmaf(p, vv.x, o[i][0]);
        o[i][1] = fmaf(p, vv.y, o[i][1]);
        o[i][2] = fmaf(p, vv.z, o[i][2]);
        o[i][3] = fmaf(p, vv.w, o[i][3]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = q0 + ty * 4 + i;
    if (t >= Tlen) continue;
    const float inv = (t < len && lrow[i] > 0.f) ? 1.0f / lrow[i] : 0.f;   // padded query rows -> 0
    T* op = obase + (long long)t * (H * HD) + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) op[j] = from_f32<T>(o[i][j] * inv);
  }
}
}  // namespace

int attention_simt(const void* qkv, int type, void* out, const long long* lens, int nb, int T, int H, cudaStream_t s) {
  dim3 grid(ceil_div(T, QT), H, nb);
  SWC_TRY(ensure_dynamic_smem((const void*)attention_simt_kernel<float>, kSmemBytes));
  SWC_TRY(ensure_dynamic_smem((const void*)attention_simt_kernel<bf16>, kSmemBytes));
  ProfScope ps(KC_ATTN, s);
  if (type == 0) attention_simt_kernel<float><<<grid, 256, kSmemBytes, s>>>((const float*)qkv, (float*)out, lens, T, H);
  else attention_simt_kernel<bf16><<<grid, 256, kSmemBytes, s>>>((const bf16*)qkv, (bf16*)out, lens, T, H);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace swc

// fp32-accumulate SIMT GEMM (CUDA-core FFMA) with the implicit-convolution tap table and the fused
// epilogues of gemm.cuh.  This is the arithmetic of the "fp32 parity mode" (indices bit-exact vs the
// reference up to near-ties) and of the two stages that must stay fp32 in every mode: the log-mel
// DFT and the inverse DFT of the iSTFT head.  128x128x16 tiles, 256 threads, 8x8 outputs per thread,
// register-staged double buffering.
#include "gemm.cuh"

namespace swc {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;

__device__ __forceinline__ int perm_col(int n) {   // keeps a thread's 8 logical columns in two float4 groups
  return ((n & 7) < 4) ? ((n >> 3) * 4 + (n & 3)) : (64 + (n >> 3) * 4 + (n & 3));
}

template <typename TA, int KIND, typename TO>
__global__ void __launch_bounds__(256, 2) gemm_simt_kernel(GemmDesc d) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int b = blockIdx.z;

  const TA* A = reinterpret_cast<const TA*>(d.A) + (long long)b * d.a_batch_stride;
  const TA* W = reinterpret_cast<const TA*>(d.W);
  const int K = d.n_taps * d.tap_k;
  const int kt_per_tap = d.tap_k / BK;
  const int num_kt = d.n_taps * kt_per_tap;

  // loader mapping: row/col = tid % 128, k-half = tid / 128
  const int lr = tid & 127, lk = (tid >> 7) * 8;
  const int wn = n0 + lr;
  const bool w_ok = wn < d.w_rows;
  const TA* wrow = W + (long long)wn * K + lk;

  float ra[8], rb[8];
  auto fetch = [&](int kt) {
    const int tap = kt / kt_per_tap, kc = (kt - tap * kt_per_tap) * BK;
    const int arow = m0 + lr + d.tap_row[tap];
    if (arow >= 0 && arow < d.a_rows) {
      load8(A + (long long)arow * d.a_row_stride + d.tap_col[tap] + kc + lk, ra);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) ra[j] = 0.0f;
    }
    if (w_ok) {
      load8(wrow + (long long)kt * BK, rb);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) rb[j] = 0.0f;
    }
  };
  auto stash = [&](int buf) {
    const int pc = perm_col(lr);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      As[buf][lk + j][lr] = ra[j];
      Bs[buf][lk + j][pc] = rb[j];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < num_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < num_kt) fetch(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], bb[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(bb) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      *reinterpret_cast<float4*>(bb + 4) = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < num_kt) stash(buf ^ 1);
    __syncthreads();
  }

  const int nc = n0 + tx * 8;
  if (nc < d.N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ty * 8 + i;
      if (m < d.m_rows)
        epi_apply<KIND, TO>(d.epi, b, m, nc, d.N, d.m_rows, acc[i], d.epi.bias ? d.epi.bias + nc : nullptr,
                            d.epi.gamma ? d.epi.gamma + nc : nullptr);
    }
  }
}

template <typename TA, int KIND, typename TO>
int launch(const GemmDesc& d, cudaStream_t s) {
  dim3 grid(ceil_div(d.N, BN), ceil_div(d.m_rows, BM), d.nb);
  ProfScope ps(KC_GEMM_SIMT, s);
  gemm_simt_kernel<TA, KIND, TO><<<grid, 256, 0, s>>>(d);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int gemm_simt(const GemmDesc& d, int kind, int a_type, int out_type, cudaStream_t s) {
  SWC_REQUIRE(d.tap_k % BK == 0 && d.n_taps >= 1 && d.n_taps <= kMaxTaps, "gemm_simt: bad tap shape (tap_k=%d taps=%d)", d.tap_k, d.n_taps);
  SWC_REQUIRE(d.m_rows > 0 && d.nb > 0 && d.N > 0, "gemm_simt: empty problem");
  SWC_REQUIRE(d.nb <= 65535 && ceil_div(d.m_rows, BM) <= 65535, "gemm_simt: grid too large");
  if (a_type == 0) {
    switch (kind) {
      case EPI_STORE: return out_type == 0 ? launch<float, EPI_STORE, float>(d, s) : launch<float, EPI_STORE, bf16>(d, s);
      case EPI_POWER: return launch<float, EPI_POWER, float>(d, s);
      case EPI_LOGMEL: return launch<float, EPI_LOGMEL, float>(d, s);
      case EPI_FSQ: return launch<float, EPI_FSQ, float>(d, s);
      case EPI_HEAD: return launch<float, EPI_HEAD, float>(d, s);
    }
  } else {
    switch (kind) {
      case EPI_STORE: return out_type == 0 ? launch<bf16, EPI_STORE, float>(d, s) : launch<bf16, EPI_STORE, bf16>(d, s);
      case EPI_FSQ: return launch<bf16, EPI_FSQ, float>(d, s);
      case EPI_HEAD: return launch<bf16, EPI_HEAD, float>(d, s);
    }
  }
  set_error("gemm_simt: unsupported kind/type combination (%d,%d,%d)", kind, a_type, out_type);
  return -1;
}

}  // namespace swc

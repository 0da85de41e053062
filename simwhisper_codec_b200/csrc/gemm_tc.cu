// bf16 tcgen05 GEMM for sm_100a: TMA-fed, TMEM-accumulated, warp-specialised, persistent.
//
//   D[b, m, n] = sum_tap sum_k A[b, m + tap_row, tap_col + k] * W[n, tap*tap_k + k]      (gemm.cuh)
//
// CTA = 128 x BN output tile (UMMA M=128, N=BN, K=16, cta_group::1), K streamed in 64-element
// (128-byte, SWIZZLE_128B) slabs through a STAGES-deep shared-memory ring:
//   warp 0   : TMA producer  (cp.async.bulk.tensor 3-D for A — the per-tap row shift is just a
//              coordinate, out-of-range rows are zero-filled by TMA = the conv zero padding —
//              and 2-D for W), arrives on full[stage] with expect_tx
//   warp 1   : MMA issuer    (one lane issues 4 tcgen05.mma per slab, tcgen05.commit frees the slot;
//              a second commit per tile publishes the TMEM accumulator)
//   warp 2   : TMEM allocator (2 accumulator buffers of BN fp32 columns, double-buffered so the
//              epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4-11: epilogue     (two warps per 32-lane TMEM group, each owning half of the BN columns:
//              per-column bias/gamma staged once per tile in shared memory, software-pipelined
//              tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue of gemm.cuh ->
//              vectorised global stores)
// Grid = min(#tiles, #SMs); tiles are walked n-fastest so concurrently resident CTAs share the A slab
// in L2 while the whole weight matrix stays L2-resident.
#include <cuda.h>
#include <cstdlib>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace swc {

namespace {

constexpr int BM = 128, BK = 64;
using namespace ptx;     // mbarrier / TMA / tcgen05 wrappers shared with gemm_tc2.cu and attention_tc.cu

struct TcParams {
  int m_rows, nb, N;
  int m_tiles, n_tiles;        // per batch
  int n_taps, kb_per_tap;
  int tap_row[kMaxTaps], tap_col[kMaxTaps];
  EpiParams epi;
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOff = STAGES * kStageBytes;
  static constexpr int kVecOff = kBarOff + (2 * STAGES + 4) * 8 + 16;          // bias[2][BN], gamma[2][BN] (fp32)
  static constexpr int kTotal = kVecOff + 4 * BN * 4 + 1024;                   // + alignment slack
  static constexpr int kEpiWarps = BN >= 64 ? 8 : 4;                           // warps 4.. ; two per TMEM lane group
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
};

template <int BN, int STAGES, int KIND, typename TO>
__global__ void __launch_bounds__(SmemLayout<BN, STAGES>::kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;     // [2] accumulator ready
  uint64_t* tempty = tfull + 2;         // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* svec = reinterpret_cast<float*>(smem + L::kVecOff);     // [2][BN] bias, then [2][BN] gamma

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;   // power of two for BN in {32,64,128,256}

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], L::kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_batch = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_per_batch * p.nb;
  const int num_kb = p.n_taps * p.kb_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_batch;
        const int r = tile - b * tiles_per_batch;
        const int m0 = (r / p.n_tiles) * BM, n0 = (r % p.n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.kb_per_tap, kc = kb - tap * p.kb_per_tap;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], L::kStageBytes);
          uint8_t* sa = smem + stage * L::kStageBytes;
          tma_load_3d<1>(&tmA, smem_u32(&full[stage]), sa, p.tap_col[tap] + kc * BK, m0 + p.tap_row[tap], b);
          tma_load_2d<1>(&tmW, smem_u32(&full[stage]), sa + L::kABytes, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + L::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (>>4) address field
            umma_bf16<1>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit<1>(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<1>(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 4 + L::kEpiWarps) {
    const int ew = warp - 4;
    const int lg = ew & 3;                   // TMEM lane group: this warp may touch lanes [32 lg, 32 lg + 32)
    const int c_base = (ew >> 2) * L::kColsPerWarp;
    const int etid = threadIdx.x - 128;
    constexpr int kEpiThreads = 32 * L::kEpiWarps;
    constexpr int kChunks = L::kColsPerWarp / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_batch;
      const int r = tile - b * tiles_per_batch;
      const int m0 = (r / p.n_tiles) * BM, n0 = (r % p.n_tiles) * BN;
      // stage this tile's per-column vectors (overlaps the MMAs of this tile)
      float* sb = svec + acc * BN;
      float* sg = svec + (2 + acc) * BN;
      for (int i = etid; i < BN; i += kEpiThreads) {
        const int n = n0 + i;
        sb[i] = (p.epi.bias && n < p.N) ? __ldg(p.epi.bias + n) : 0.0f;
        sg[i] = (p.epi.gamma && n < p.N) ? __ldg(p.epi.gamma + n) : 1.0f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int m = m0 + lg * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + c_base;
      uint32_t rr[2][32];
      tmem_ld32(taddr, rr[0]);
#pragma unroll
      for (int cb = 0; cb < kChunks; ++cb) {
        tmem_ld_wait();
        if (cb + 1 < kChunks) tmem_ld32(taddr + (cb + 1) * 32, rr[(cb + 1) & 1]);
        if (m < p.m_rows) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cl = c_base + cb * 32 + j * 8;
            const int nc = n0 + cl;
            if (nc < p.N) {
              float v[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(rr[cb & 1][j * 8 + q]);
              epi_apply<KIND, TO>(p.epi, b, m, nc, p.N, p.m_rows, v, p.epi.bias ? sb + cl : nullptr,
                                  p.epi.gamma ? sg + cl : nullptr);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

// ---- host side -------------------------------------------------------------------------------------
template <int BN, int STAGES, int KIND, typename TO>
int launch(const GemmDesc& d, int num_sms, cudaStream_t s) {
  using L = SmemLayout<BN, STAGES>;
  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[3] = {(cuuint64_t)d.a_cols, (cuuint64_t)d.a_rows, (cuuint64_t)d.nb};
    cuuint64_t strides[2] = {(cuuint64_t)d.a_row_stride * 2, (cuuint64_t)(d.nb > 1 ? d.a_batch_stride : (long long)d.a_row_stride * d.a_rows) * 2};
    cuuint32_t box[3] = {BK, BM, 1};
    SWC_TRY(make_tmap(&tmA, 1, d.A, 3, dims, strides, box));
  }
  {
    const int K = d.n_taps * d.tap_k;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)d.w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {BK, BN};
    SWC_TRY(make_tmap(&tmW, 1, d.W, 2, dims, strides, box));
  }
  TcParams p{};
  p.m_rows = d.m_rows; p.nb = d.nb; p.N = d.N;
  p.m_tiles = ceil_div(d.m_rows, BM);
  p.n_tiles = ceil_div(d.N, BN);
  p.n_taps = d.n_taps; p.kb_per_tap = d.tap_k / BK;
  for (int i = 0; i < d.n_taps; ++i) { p.tap_row[i] = d.tap_row[i]; p.tap_col[i] = d.tap_col[i]; }
  p.epi = d.epi;
  auto kern = gemm_tc_kernel<BN, STAGES, KIND, TO>;
  SWC_TRY(ensure_dynamic_smem((const void*)kern, L::kTotal));
  const long long total = (long long)p.m_tiles * p.n_tiles * p.nb;
  const int grid = (int)std::min<long long>(total, num_sms);
  ProfScope ps(KC_GEMM_TC, s);
  kern<<<grid, L::kThreads, L::kTotal, s>>>(tmA, tmW, p);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int KIND, typename TO>
int launch_bn(const GemmDesc& d, int num_sms, cudaStream_t s) {
  if (d.N <= 32) return launch<32, 8, KIND, TO>(d, num_sms, s);
  if (d.N <= 128 || d.N % 256 != 0) return launch<128, 6, KIND, TO>(d, num_sms, s);
  return launch<256, 4, KIND, TO>(d, num_sms, s);
}

}  // namespace

static int g_gemm_variant = -1;
void set_gemm_variant(int v) { g_gemm_variant = v; }
int get_gemm_variant() {
  if (g_gemm_variant < 0) {
    const char* e = getenv("SWC_GEMM_VARIANT");
    g_gemm_variant = e ? atoi(e) : 2;
  }
  return g_gemm_variant;
}

int gemm_tc(const GemmDesc& d, int kind, int out_type, int num_sms, cudaStream_t s) {
  if (kind == EPI_STORE && get_gemm_variant() > 0 && gemm_tc2_eligible(d) && !(d.epi.residual && out_type != 0) && !(d.epi.act == 1 && out_type != 0 && !d.epi.out_planes)) return gemm_tc2(d, out_type, num_sms, get_gemm_variant(), s);
  SWC_REQUIRE(!d.epi.out_planes, "gemm_tc: a planes output is only written by the pair kernel (problem not eligible for it)");
  SWC_REQUIRE(d.tap_k % BK == 0 && d.n_taps >= 1 && d.n_taps <= kMaxTaps, "gemm_tc: tap_k=%d must be a multiple of 64 (taps=%d)", d.tap_k, d.n_taps);
  SWC_REQUIRE(d.m_rows > 0 && d.nb > 0 && d.N > 0, "gemm_tc: empty problem");
  SWC_REQUIRE(((uintptr_t)d.A & 15) == 0 && ((uintptr_t)d.W & 15) == 0, "gemm_tc: operands must be 16-byte aligned");
  SWC_REQUIRE((d.a_row_stride * 2) % 16 == 0 && (d.a_batch_stride * 2) % 16 == 0, "gemm_tc: operand strides must be multiples of 16 bytes");
  switch (kind) {
    case EPI_STORE: return out_type == 0 ? launch_bn<EPI_STORE, float>(d, num_sms, s) : launch_bn<EPI_STORE, bf16>(d, num_sms, s);
    case EPI_FSQ: return launch_bn<EPI_FSQ, float>(d, num_sms, s);
    case EPI_POWER: return launch_bn<EPI_POWER, float>(d, num_sms, s);
    case EPI_HEAD: return launch_bn<EPI_HEAD, float>(d, num_sms, s);
  }
  set_error("gemm_tc: unsupported epilogue kind %d", kind);
  return -1;
}

}  // namespace swc

// bf16 tcgen05 GEMM, second generation: CTA-pair MMA (cta_group::2) and TMA-store epilogue.
//
//   D[b, m, n] (+)= act(sum_tap sum_k A[b, m + tap_row, tap_col + k] * W[n, tap*tap_k + k] + bias[n]) * gamma[n]
//
// One cluster of CG CTAs (CG = 2: the two SMs of a TPC) owns a (128*CG) x 256 output tile.  Each CTA stages its
// own 128 rows of A and its 256/CG rows of W per 64-wide K slab, so with CG = 2 the pair reads every W slab
// from L2 once instead of twice (the cta_group::1 kernel of gemm_tc.cu is L2->SM bandwidth bound: 96 B/clk/SM
// at full tensor rate; the pair needs 64).  UMMA M = 128*CG, N = 256, K = 16; the fp32 accumulator
// (128 lanes x 256 columns per CTA) is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1.
//   warp 0    : TMA producer (both CTAs; with CG = 2 all completions land on the leader's `full` barrier)
//   warp 1    : MMA issuer (leader CTA only); tcgen05.commit multicasts `empty`/`tfull` to both CTAs
//   warp 2    : TMEM allocator
//   warps 4-11: epilogue: tcgen05.ld 32 lanes x 32 columns -> bias / GELU / gamma (packed f32x2) -> 128B-swizzled
//               shared staging (one 32-row x 128-byte box per warp) -> cp.async.bulk.tensor store.  Stores are
//               whole 128-byte lines and TMA clips rows >= m_rows and columns >= N.
// Persistent: grid = CG * min(#tiles, #SMs / CG); tiles are walked n-fastest so the clusters resident at any
// time share A rows in L2.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <set>
#include <type_traits>
#include <utility>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace swc {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device, per-function setting: remember the (device, function)
// pairs that have it instead of one process-wide flag, so a second GPU in the same process gets its opt-in too.
int ensure_dynamic_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  SWC_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, func})) return 0;
  SWC_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({dev, func});
  return 0;
}

EncodeTiledFn tmap_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap(CUtensorMap* map, int dtype, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box) {
  EncodeTiledFn fn = tmap_encode_fn();
  SWC_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t elem[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                  const_cast<void*>(ptr), dims, strides_bytes, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SWC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}

namespace {

using namespace ptx;

constexpr int BM = 128, BN = 256, BK = 64;

struct Tc2Params {
  int m_rows, nb, N;
  int m_tiles, n_tiles;        // per batch; an m tile spans 128*CG rows
  int n_taps, kb_per_tap;
  int tap_row[kMaxTaps], tap_col[kMaxTaps];
  const float* bias;
  const float* gamma;
};

template <int CG, int STAGES, int NBUF, int EW>
struct Smem2 {
  static constexpr int kEpiWarps = EW;                                  // 8 or 16: four TMEM lane groups x EW/4 column slices
  static constexpr int kThreads = 128 + 32 * EW;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / CG) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutOff = STAGES * kStageBytes;
  static constexpr int kOutBytesPerWarp = 32 * 128;                     // 32 rows x 128-byte swizzle line
  static constexpr int kBarOff = kOutOff + kEpiWarps * NBUF * kOutBytesPerWarp;
  static constexpr int kTotal = kBarOff + (2 * STAGES + 4) * 8 + 16 + 1024;   // + alignment slack
  static_assert(kTotal <= 232448, "shared memory budget");
};

template <int ACT>
__device__ __forceinline__ float act_apply(float x) {
  if constexpr (ACT == 2) return gelu_fast(x);
  else if constexpr (ACT == 1) return gelu_erf(x);
  else return x;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ unsigned long long pack_f32x2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
// 4 accumulator columns -> (act(acc + bias)) * gamma.  bias / gamma are read straight from global memory: the
// address is uniform across the warp (one L1 transaction per load) and the vectors stay L1-resident.  FAST: the whole
// 32-column chunk lies inside N and a bias exists, so no per-chunk predicates (the common case; the general form costs
// a compare, a select and a register clear per four columns).
template <int ACT, bool GAMMA, bool FAST>
__device__ __forceinline__ void epi4(const uint32_t* acc, const float* sb, const float* sg, bool in_range, float* v) {
  float4 b4;
  if constexpr (FAST) b4 = __ldg(reinterpret_cast<const float4*>(sb));
  else b4 = (sb && in_range) ? __ldg(reinterpret_cast<const float4*>(sb)) : make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long lo, hi;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(lo) : "l"(pack_f32x2(__uint_as_float(acc[0]), __uint_as_float(acc[1]))), "l"(pack_f32x2(b4.x, b4.y)));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(hi) : "l"(pack_f32x2(__uint_as_float(acc[2]), __uint_as_float(acc[3]))), "l"(pack_f32x2(b4.z, b4.w)));
  if constexpr (ACT == 2) {
    lo = gelu_fast2(lo);
    hi = gelu_fast2(hi);
  }
  if constexpr (GAMMA) {
    float4 g4;
    if constexpr (FAST) g4 = __ldg(reinterpret_cast<const float4*>(sg));
    else g4 = in_range ? __ldg(reinterpret_cast<const float4*>(sg)) : make_float4(1.f, 1.f, 1.f, 1.f);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(lo) : "l"(lo), "l"(pack_f32x2(g4.x, g4.y)));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hi) : "l"(hi), "l"(pack_f32x2(g4.z, g4.w)));
  }
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v[0]), "=f"(v[1]) : "l"(lo));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2]), "=f"(v[3]) : "l"(hi));
  if constexpr (ACT == 1) {
    v[0] = gelu_erf(v[0]); v[1] = gelu_erf(v[1]); v[2] = gelu_erf(v[2]); v[3] = gelu_erf(v[3]);
  }
}

// 32 accumulator columns of one row -> epilogue math -> the row's 128-byte line of the staging box (16-byte chunks
// XOR-swizzled by the row).  bf16 results fill half a line per call (`half` selects which).
template <int ACT, bool GAMMA, bool FAST, bool kF32>
__device__ __forceinline__ void convert_sub(const uint32_t (&acc)[32], const float* sbc, const float* sgc, int ncol, int N,
                                            uint32_t box, uint32_t sw, int half) {
  if constexpr (kF32) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v[4];
      epi4<ACT, GAMMA, FAST>(&acc[4 * q], sbc ? sbc + 4 * q : nullptr, sgc + 4 * q, ncol + 4 * q + 4 <= N, v);
      st_shared_v4(box + (((uint32_t)q ^ sw) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                   __float_as_uint(v[3]));
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
      epi4<ACT, GAMMA, FAST>(&acc[8 * q], sbc ? sbc + 8 * q : nullptr, sgc + 8 * q, ncol + 8 * q + 4 <= N, v);
      epi4<ACT, GAMMA, FAST>(&acc[8 * q + 4], sbc ? sbc + 8 * q + 4 : nullptr, sgc + 8 * q + 4, ncol + 8 * q + 8 <= N, v + 4);
      const uint32_t chunk = (uint32_t)(half * 4 + q);
      st_shared_v4(box + ((chunk ^ sw) << 4), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                   pack_bf16(v[6], v[7]));
    }
  }
}

// One epilogue warp's share of an output tile: its 32 TMEM lanes x kSlice accumulator columns -> bias / activation /
// gamma -> 128B-swizzled staging boxes -> TMA store (or reduce-add).  `taddr` addresses the warp's first column of the
// accumulator buffer; the buffer is handed back (arrive on `tempty_addr`) as soon as it is in registers.
template <int NBUF, int EW, int ACT, bool GAMMA, bool REDUCE, typename TO, bool PLANES = false>
__device__ __forceinline__ void epilogue_tile(const CUtensorMap* tmO, const Tc2Params& p, uint32_t taddr, uint32_t tempty_addr,
                                              uint8_t* stage_out, int& obuf, int lane, int lg, int c_base, int b, int m0, int n0) {
  constexpr int kSlice = BN / (EW / 4);
  constexpr bool kF32 = sizeof(TO) == 4;
  constexpr int kOutBytesPerWarp = 32 * 128;
  const uint32_t sw = (uint32_t)(lane & 7);
  uint32_t rr[2][32];
  tmem_ld32(taddr, rr[0]);
#pragma unroll
  for (int sub = 0; sub < kSlice / 32; ++sub) {
    tmem_ld_wait();
    if (sub + 1 < kSlice / 32) tmem_ld32(taddr + (sub + 1) * 32, rr[(sub + 1) & 1]);
    if (sub == kSlice / 32 - 1) {                       // accumulator fully in registers: hand the TMEM buffer back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
    }
    if constexpr (PLANES) {
      // bf16x3 hand-over: the fp32 result goes out as two bf16 planes (hi | lo), one staging box each; two 64-column
      // TMA stores per pair of chunks, at columns n and N + n of the 2N-wide output rows
      static_assert(!PLANES || (NBUF == 2 && sizeof(TO) == 2 && !REDUCE && !GAMMA), "planes epilogue: bf16, two staging boxes");
      if ((sub & 1) == 0) {
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      const uint32_t box_hi = smem_u32(stage_out) + (uint32_t)lane * 128u, box_lo = box_hi + kOutBytesPerWarp;
      const int ncol = n0 + c_base + sub * 32;
      const float* sbc = p.bias ? p.bias + ncol : nullptr;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[8];
        if (sbc != nullptr && ncol + 32 <= p.N) {
          epi4<ACT, false, true>(&rr[sub & 1][8 * q], sbc + 8 * q, nullptr, true, v);
          epi4<ACT, false, true>(&rr[sub & 1][8 * q + 4], sbc + 8 * q + 4, nullptr, true, v + 4);
        } else {
          epi4<ACT, false, false>(&rr[sub & 1][8 * q], sbc ? sbc + 8 * q : nullptr, nullptr, ncol + 8 * q + 4 <= p.N, v);
          epi4<ACT, false, false>(&rr[sub & 1][8 * q + 4], sbc ? sbc + 8 * q + 4 : nullptr, nullptr, ncol + 8 * q + 8 <= p.N, v + 4);
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          const float2 hf = __bfloat1622float2(h2);
          hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[e] = pack_bf16(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        }
        const uint32_t chunk = (uint32_t)((sub & 1) * 4 + q);
        st_shared_v4(box_hi + ((chunk ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
        st_shared_v4(box_lo + ((chunk ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
      }
      if ((sub & 1) == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int col = n0 + c_base + (sub >> 1) * 64;
          tma_store_3d(tmO, stage_out, col, m0 + lg * 32, b);
          tma_store_3d(tmO, stage_out + kOutBytesPerWarp, p.N + col, m0 + lg * 32, b);
          tma_store_commit();
        }
      }
      continue;
    }
    const bool new_box = kF32 || (sub & 1) == 0;
    if (new_box) {                        // the staging box must have been read by its previous TMA store
      if (lane == 0) tma_store_wait_read<NBUF - 1>();
      __syncwarp();
    }
    const uint32_t box = smem_u32(stage_out + obuf * kOutBytesPerWarp) + (uint32_t)lane * 128u;
    const int ncol = n0 + c_base + sub * 32;      // first of this sub-chunk's 32 columns
    const float* sbc = p.bias ? p.bias + ncol : nullptr;
    const float* sgc = GAMMA ? p.gamma + ncol : nullptr;
    if (sbc != nullptr && ncol + 32 <= p.N) convert_sub<ACT, GAMMA, true, kF32>(rr[sub & 1], sbc, sgc, ncol, p.N, box, sw, sub & 1);
    else convert_sub<ACT, GAMMA, false, kF32>(rr[sub & 1], sbc, sgc, ncol, p.N, box, sw, sub & 1);
    const bool box_done = kF32 || (sub & 1) == 1;
    if (box_done) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int col = n0 + c_base + (kF32 ? sub * 32 : (sub >> 1) * 64);
        // REDUCE: the residual add out += tile is done by the L2 (one tile per output element: deterministic)
        if constexpr (REDUCE) tma_reduce_add_3d(tmO, stage_out + obuf * kOutBytesPerWarp, col, m0 + lg * 32, b);
        else tma_store_3d(tmO, stage_out + obuf * kOutBytesPerWarp, col, m0 + lg * 32, b);
        tma_store_commit();
      }
      obuf = (obuf + 1 == NBUF) ? 0 : obuf + 1;
    }
  }
}

template <int CG, int STAGES, int NBUF, int EW, int ACT, bool GAMMA, bool REDUCE, typename TO, bool PLANES = false>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmO, const Tc2Params p) {
  using L = Smem2<CG, STAGES, NBUF, EW>;
  constexpr int kEpiWarps = EW;
  constexpr int kSlice = BN / (EW / 4);          // accumulator columns per epilogue warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;     // [2] accumulator ready (per CTA)
  uint64_t* tempty = tfull + 2;         // [2] accumulator drained (leader CTA collects both CTAs' epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
  const bool leader = rank == 0;
  constexpr uint32_t kTmemCols = 2 * BN;   // 512: the whole TMEM, one CTA per SM

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CG * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if constexpr (CG > 1) cluster_sync();      // peer barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  const int cid = blockIdx.x / CG, ncl = gridDim.x / CG;
  const int tiles_per_batch = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_per_batch * p.nb;
  const int num_kb = p.n_taps * p.kb_per_tap;

  if (warp == 0) {
    // whole warp runs the loop (uniform control flow); one elected lane issues the TMA instructions
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl) {
      const int b = tile / tiles_per_batch;
      const int r = tile - b * tiles_per_batch;
      const int m0 = (r / p.n_tiles) * (BM * CG) + (int)rank * BM;
      const int n0 = (r % p.n_tiles) * BN + (int)rank * (BN / CG);
      int tap = 0, kc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          // all bytes of the pair land on the leader's barrier; the leader alone arrives (expect_tx of both halves)
          const uint32_t bar = CG == 1 ? smem_u32(&full[stage]) : mapa(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], CG * L::kStageBytes);
          uint8_t* sa = smem + stage * L::kStageBytes;
          tma_load_3d<CG>(&tmA, bar, sa, p.tap_col[tap] + kc * BK, m0 + p.tap_row[tap], b);
          tma_load_2d<CG>(&tmW, bar, sa + L::kABytes, kb * BK, n0);
        }
        __syncwarp();
        if (++kc == p.kb_per_tap) { kc = 0; ++tap; }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = make_idesc(BM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + L::kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (>>4) address field
              umma_bf16<CG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit<CG>(&empty[stage]);
            if (kb == num_kb - 1) umma_commit<CG>(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int lg = ew & 3;                    // TMEM lane group of this warp: lanes [32 lg, 32 lg + 32)
    const int c_base = (ew >> 2) * kSlice;    // this warp's accumulator columns
    uint8_t* stage_out = smem + L::kOutOff + ew * NBUF * L::kOutBytesPerWarp;
    const uint32_t tempty_leader[2] = {CG == 1 ? smem_u32(&tempty[0]) : mapa(smem_u32(&tempty[0]), 0),
                                       CG == 1 ? smem_u32(&tempty[1]) : mapa(smem_u32(&tempty[1]), 0)};
    int acc = 0, obuf = 0;
    uint32_t acc_phase = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl) {
      const int b = tile / tiles_per_batch;
      const int r = tile - b * tiles_per_batch;
      const int m0 = (r / p.n_tiles) * (BM * CG) + (int)rank * BM;
      const int n0 = (r % p.n_tiles) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<NBUF, EW, ACT, GAMMA, REDUCE, TO, PLANES>(&tmO, p, tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + c_base,
                                                      tempty_leader[acc], stage_out, obuf, lane, lg, c_base, b, m0, n0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG > 1) cluster_sync();      // the peer may still be arriving on / reading from this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

template <int CG, int STAGES, int NBUF, int EW, int ACT, bool GAMMA, bool REDUCE, typename TO, bool PLANES = false>
int launch2(const GemmDesc& d, int num_sms, cudaStream_t s) {
  using L = Smem2<CG, STAGES, NBUF, EW>;
  constexpr int out_type = sizeof(TO) == 4 ? 0 : 1;
  CUtensorMap tmA, tmW, tmO;
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
    cuuint32_t box[2] = {BK, BN / CG};
    SWC_TRY(make_tmap(&tmW, 1, d.W, 2, dims, strides, box));
  }
  {
    const size_t es = sizeof(TO);
    const long long row_stride = d.epi.out_row_stride * d.epi.out_row_mul;
    const char* base = (const char*)d.epi.out + (size_t)d.epi.out_row_off * d.epi.out_row_stride * es;
    cuuint64_t dims[3] = {(cuuint64_t)(PLANES ? 2 * d.N : d.N), (cuuint64_t)d.m_rows, (cuuint64_t)d.nb};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride * es, (cuuint64_t)(d.nb > 1 ? d.epi.out_batch_stride : row_stride * d.m_rows) * es};
    cuuint32_t box[3] = {(cuuint32_t)(128 / es), 32, 1};
    SWC_TRY(make_tmap(&tmO, out_type, base, 3, dims, strides, box));
  }
  Tc2Params p{};
  p.m_rows = d.m_rows; p.nb = d.nb; p.N = d.N;
  p.m_tiles = ceil_div(d.m_rows, BM * CG);
  p.n_tiles = ceil_div(d.N, BN);
  p.n_taps = d.n_taps; p.kb_per_tap = d.tap_k / BK;
  for (int i = 0; i < d.n_taps; ++i) { p.tap_row[i] = d.tap_row[i]; p.tap_col[i] = d.tap_col[i]; }
  p.bias = d.epi.bias; p.gamma = d.epi.gamma;
  auto kern = gemm_tc2_kernel<CG, STAGES, NBUF, EW, ACT, GAMMA, REDUCE, TO, PLANES>;
  SWC_TRY(ensure_dynamic_smem((const void*)kern, L::kTotal));
  const long long total = (long long)p.m_tiles * p.n_tiles * p.nb;
  const int grid = CG * (int)std::min<long long>(total, num_sms / CG);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(L::kThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  ProfScope ps(KC_GEMM_TC, s);
  SWC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, tmO, p));
  return 0;
}

template <int CG, int STAGES, int NBUF, int EW>
int dispatch2(const GemmDesc& d, int out_type, int num_sms, cudaStream_t s) {
  const int act = d.epi.act;
  const bool gamma = d.epi.gamma != nullptr;
  const bool reduce = d.epi.residual != nullptr;      // eligibility guarantees residual == out (fp32, same addressing)
  if (out_type == 0) {
    if (reduce) {
      if (act == 0 && !gamma) return launch2<CG, STAGES, NBUF, EW, 0, false, true, float>(d, num_sms, s);
      if (act == 0 && gamma) return launch2<CG, STAGES, NBUF, EW, 0, true, true, float>(d, num_sms, s);
    } else {
      if (act == 0 && !gamma) return launch2<CG, STAGES, NBUF, EW, 0, false, false, float>(d, num_sms, s);
      if (act == 0 && gamma) return launch2<CG, STAGES, NBUF, EW, 0, true, false, float>(d, num_sms, s);
      if (act == 2 && !gamma) return launch2<CG, STAGES, NBUF, EW, 2, false, false, float>(d, num_sms, s);
      if (act == 1 && !gamma) return launch2<CG, STAGES, NBUF, EW, 1, false, false, float>(d, num_sms, s);   // exact erf (bf16x3 mode)
    }
  } else if (!reduce) {
    if (act == 0 && !gamma) return launch2<CG, STAGES, NBUF, EW, 0, false, false, bf16>(d, num_sms, s);
    if (act == 2 && !gamma) return launch2<CG, STAGES, NBUF, EW, 2, false, false, bf16>(d, num_sms, s);
  }
  set_error("gemm_tc2: unsupported epilogue (act %d, gamma %d, out type %d)", act, (int)gamma, out_type);
  return -1;
}

}  // namespace

// which EPI_STORE problems the second-generation kernel takes
bool gemm_tc2_eligible(const GemmDesc& d) {
  const EpiParams& e = d.epi;
  if (e.gamma && e.act != 0) return false;       // gamma only accompanies the plain fp32 pwconv2 epilogue
  // a residual is taken only as an in-place accumulation out += tile (TMA reduce-add), never as a separate operand
  if (e.residual && (e.residual != e.out || e.act != 0 || e.res_row_stride != e.out_row_stride * e.out_row_mul ||
                     e.out_row_off != 0 || (d.nb > 1 && e.res_batch_stride != e.out_batch_stride)))
    return false;
  return e.out2 == nullptr && (e.act == 0 || e.act == 2 || (e.act == 1 && !e.residual)) && d.N >= 256 && d.N % 8 == 0 &&
         d.tap_k % BK == 0 && ((uintptr_t)e.out & 15) == 0;
}

// variant: 1 = single-CTA tiles + TMA store, 2 = CTA pairs, 3 = CTA pairs with 16 epilogue warps
int gemm_tc2(const GemmDesc& d, int out_type, int num_sms, int variant, cudaStream_t s) {
  SWC_REQUIRE(gemm_tc2_eligible(d), "gemm_tc2: problem not eligible (residual/out2/act/N)");
  SWC_REQUIRE(d.m_rows > 0 && d.nb > 0, "gemm_tc2: empty problem");
  SWC_REQUIRE(((uintptr_t)d.A & 15) == 0 && ((uintptr_t)d.W & 15) == 0, "gemm_tc2: operands must be 16-byte aligned");
  if (d.epi.out_planes) {     // bf16x3 hand-over of an exact-erf GELU (or plain) result to the next three-product GEMM
    SWC_REQUIRE(out_type == 1 && d.N % 64 == 0 && !d.epi.gamma && !d.epi.residual && (d.epi.act == 0 || d.epi.act == 1),
                "gemm_tc2: planes output needs a bf16 result, N %% 64 == 0 and no gamma / residual");
    if (d.epi.act == 1) return launch2<2, 5, 2, 8, 1, false, false, bf16, true>(d, num_sms, s);
    return launch2<2, 5, 2, 8, 0, false, false, bf16, true>(d, num_sms, s);
  }
  if (variant == 1) return dispatch2<1, 4, 1, 8>(d, out_type, num_sms, s);
  if (variant == 3) return dispatch2<2, 5, 1, 16>(d, out_type, num_sms, s);  // 16 epilogue warps (64 columns each), 5-stage ring
  // Short K with a bf16 result (pwconv1, to_stacked: 8 K slabs per tile, two staging boxes per warp and tile): the tile
  // is epilogue-paced, so a second staging box per warp (no wait for the previous TMA store in mid-tile) is worth more
  // than the sixth operand stage (+2.7 % on pwconv1; the long-K fp32 shapes lose 5 % with it and keep six stages).
  if (out_type == 1 && d.n_taps * d.tap_k <= 8 * BK) return dispatch2<2, 5, 2, 8>(d, out_type, num_sms, s);
  return dispatch2<2, 6, 1, 8>(d, out_type, num_sms, s);                     // deepest operand ring that fits 227 KB
}

}  // namespace swc

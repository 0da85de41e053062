// fp32-class flash attention on tcgen05 for the bf16x3 precision (reference audiocodec/nn/modules.py:145-187): q, k, v arrive
// as (hi | lo) bf16 planes of the fp32 rows (x = hi + lo up to 2^-17, written by the qkv GEMM's planes epilogue) and both
// contractions run as three bf16 products with fp32 accumulation in TMEM:
//
//     S = Q_h K_h^T + Q_l K_h^T + Q_h K_l^T            (12 UMMA M128 N128 K16 per key block, operands from shared memory)
//     O += P_h V_h + P_l V_h + P_h V_l                 (24 UMMA M128 N64 K16, A = P planes from TMEM, B = V MN-major)
//
// Same persistent, warp-specialised structure as attention_tc.cu (work item = (batch, head, two 128-query tiles); warp 0 TMA
// producer, warps 1 / 2 MMA issuers of tile 0 / 1, warps 4-19 softmax, two threads per query row), with these differences:
//   * the probabilities are split in registers, p = p_h + p_l, and written as two bf16 planes INTO THE COLUMNS OF S
//     (S_i: 128 fp32 columns = P_h 64 + P_l 64 packed columns), so S_i(j+1) is issued behind P_i(j) V(j) by the same thread
//     (the tensor pipe executes one thread's MMAs in order): no separate "S free" / "P V done" barriers inside an item, the
//     arrival of S_i(j+1) implies that P_i(j) V has retired and O_i is stable for the lazy rescale;
//   * Q tiles are single-buffered (four 16 KB planes), the K | V ring has two stages of four planes (64 KB each);
//   * the normalised output goes out as (hi | lo) bf16 planes, the operand of the three-product out_proj GEMM.
// TMEM columns: S0 / P0 0-127 | S1 / P1 128-255 | O0 256-319 | O1 320-383.
#include <algorithm>
#include <cstdlib>

#include "attention_common.cuh"

namespace swc {

namespace {

using namespace ptx;
using namespace attn;

constexpr int STAGES = 2;
constexpr int kQOff = 0;                                   // Q: [tile][plane]
constexpr int kKVOff = 4 * kTileBytes;                     // ring: [stage][K_h, K_l, V_h, V_l]
constexpr int kStageBytes = 4 * kTileBytes;
constexpr int kBarOff = kKVOff + STAGES * kStageBytes;
constexpr int kNumBars = 2 + 2 + 2 * STAGES + 2 + 2 + 2;
constexpr int kXchOff = kBarOff + kNumBars * 8 + 16;       // float [tile][buffer][part][row] exchange slots
constexpr int kSmemBytes = kXchOff + 2 * 2 * 2 * QT * 4 + 1024;
constexpr uint32_t kColS = 0, kColO = 256;
constexpr int SPLIT = 2;

// p = hi + lo with hi = bf16(p), lo = bf16(p - hi): two values per call, packed
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack2(a - hf.x, b - hf.y);
}

template <typename TAB>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_x3_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p, const __grid_constant__ TAB tab) {
  constexpr bool kRagged = std::is_same<TAB, RaggedTable>::value;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* q_full = bars;                 // [tile]
  uint64_t* q_empty = q_full + 2;          // [tile] last S of the item issued and retired
  uint64_t* kv_full = q_empty + 2;         // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;   // [STAGES] both tiles' P V of the block retired
  uint64_t* s_full = kv_empty + STAGES;    // [tile] S_i written (and every earlier MMA of the tile retired)
  uint64_t* p_full = s_full + 2;           // [tile] P_i planes (and any O_i rescale) written by the 8 softmax warps
  uint64_t* o_done = p_full + 2;           // [tile] last P V of the item retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4 * SPLIT); mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  const int D3 = 3 * p.H * HD;               // columns of one plane of a qkv row

  if (warp < 4) setmaxnreg_dec<kRegsIssue>();
  else setmaxnreg_inc<kRegsSoftmax>();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t q_cnt[2] = {0, 0};
    uint32_t kv_it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item(p, tab, item);
      if (it.dead) continue;
      for (int i = 0; i < it.n_act; ++i) {
        mbar_wait(&q_empty[i], (q_cnt[i] & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&q_full[i], 2 * kTileBytes);
          uint8_t* dst = smem + kQOff + i * 2 * kTileBytes;
          tma_load_2d<1>(&tmQKV, smem_u32(&q_full[i]), dst, it.h * HD, it.row0 + it.q0 + i * QT);
          tma_load_2d<1>(&tmQKV, smem_u32(&q_full[i]), dst + kTileBytes, D3 + it.h * HD, it.row0 + it.q0 + i * QT);
        }
        __syncwarp();
        ++q_cnt[i];
      }
      for (int j = 0; j < it.n_kt; ++j, ++kv_it) {
        const uint32_t st = kv_it % STAGES, ph = (kv_it / STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[st], kStageBytes);
          uint8_t* dst = smem + kKVOff + st * kStageBytes;
          const int r = it.row0 + j * KT, ck = (p.H + it.h) * HD, cv = (2 * p.H + it.h) * HD;
          tma_load_2d<1>(&tmQKV, smem_u32(&kv_full[st]), dst, ck, r);
          tma_load_2d<1>(&tmQKV, smem_u32(&kv_full[st]), dst + kTileBytes, D3 + ck, r);
          tma_load_2d<1>(&tmQKV, smem_u32(&kv_full[st]), dst + 2 * kTileBytes, cv, r);
          tma_load_2d<1>(&tmQKV, smem_u32(&kv_full[st]), dst + 3 * kTileBytes, D3 + cv, r);
        }
        __syncwarp();
      }
    }
  } else if (warp <= 2) {
    // ------------------------------------------------------------------ MMA issuer of query tile i
    const int i = warp - 1;
    constexpr uint32_t idesc_s = make_idesc(QT, KT);          // Q K^T: both K-major
    constexpr uint32_t idesc_o = make_idesc(QT, HD, 1);       // P V: B = V is MN-major ([key][dim], dim contiguous)
    const uint32_t s_tmem = tmem_base + kColS + i * 128, o_tmem = tmem_base + kColO + i * 64;
    const uint32_t q_smem = smem_u32(smem + kQOff + i * 2 * kTileBytes);
    uint32_t g = 0;             // ring index of the current (item, key block), in the producer's order
    uint32_t n_q = 0, n_p = 0;  // items / key blocks this tile has taken part in
    auto issue_s = [&](uint32_t ring) {      // S_i = Q_h K_h^T + Q_l K_h^T + Q_h K_l^T of the key block in ring entry `ring`
      const uint32_t slot = ring % STAGES;
      mbar_wait(&kv_full[slot], (ring / STAGES) & 1);
      tc_fence_after();
      const uint32_t k_smem = smem_u32(smem + kKVOff + slot * kStageBytes);
      const uint64_t qh = make_smem_desc(q_smem), ql = make_smem_desc(q_smem + kTileBytes);
      const uint64_t kh = make_smem_desc(k_smem), kl = make_smem_desc(k_smem + kTileBytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16<1>(s_tmem, qh + 2 * k, kh + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16<1>(s_tmem, ql + 2 * k, kh + 2 * k, idesc_s, 1);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16<1>(s_tmem, qh + 2 * k, kl + 2 * k, idesc_s, 1);
        umma_commit<1>(&s_full[i]);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item(p, tab, item);
      if (it.dead) continue;
      if (i >= it.n_act) {
        // this tile sits the item out but still owes the ring its releases: in ring order, once each slot has landed
        for (int j = 0; j < it.n_kt; ++j, ++g) {
          mbar_wait(&kv_full[g % STAGES], (g / STAGES) & 1);
          if (elect_one()) mbar_arrive(&kv_empty[g % STAGES]);
          __syncwarp();
        }
        continue;
      }
      mbar_wait(&q_full[i], n_q & 1);
      ++n_q;
      issue_s(g);
      for (int j = 0; j < it.n_kt; ++j, ++g) {
        const uint32_t slot = g % STAGES;
        mbar_wait(&p_full[i], n_p & 1);
        ++n_p;
        tc_fence_after();
        const uint32_t v_smem = smem_u32(smem + kKVOff + slot * kStageBytes + 2 * kTileBytes);
        const uint64_t vh = make_smem_desc(v_smem), vl = make_smem_desc(v_smem + kTileBytes);
        const bool last = j + 1 == it.n_kt;
        if (elect_one()) {
          // A: 16 keys = 8 packed columns of a P plane; B: 16 key rows of V = 2 swizzle atoms of 8 rows x 128 B (+128 in >>4 units)
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) umma_bf16_ts(o_tmem, s_tmem + k * 8, vh + (uint64_t)(k * 128), idesc_o, (j | k) != 0);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) umma_bf16_ts(o_tmem, s_tmem + 64 + k * 8, vh + (uint64_t)(k * 128), idesc_o, 1);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) umma_bf16_ts(o_tmem, s_tmem + k * 8, vl + (uint64_t)(k * 128), idesc_o, 1);
          umma_commit<1>(&kv_empty[slot]);
          if (last) {
            umma_commit<1>(&o_done[i]);
            umma_commit<1>(&q_empty[i]);                        // every S of this item has retired as well
          }
        }
        __syncwarp();
        if (!last) issue_s(g + 1);                              // overwrites the P planes just consumed: same thread, in order
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax warps
    constexpr int NC = KT / SPLIT, ND = HD / SPLIT;
    const int sw = warp - 4;                          // softmax warp index
    const int i = sw / (4 * SPLIT);                   // query tile
    const int part = (sw >> 2) % SPLIT;               // column part of the row
    const int lg = warp & 3;                          // TMEM lane group this warp may access
    const int pair_bar = 1 + i * 4 + lg;              // the two warps that share 32 rows
    const int row = lg * 32 + lane;                   // query row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    const uint32_t s_addr = lane_addr + kColS + i * 128, o_addr = lane_addr + kColO + i * 64 + part * ND;
    const int DO = p.H * HD;
    const uint32_t xch = smem_u32(smem + kXchOff) + (uint32_t)(i * (2 * SPLIT * QT) + row) * 4u;   // [2 buffers][SPLIT][128 rows]
    uint32_t s_cnt = 0, o_cnt = 0, x_cnt = 0;
    long long len_next = blockIdx.x < p.n_items ? item_len_raw(p, tab, blockIdx.x) : 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item(p, tab, item, len_next);
      if (item + (int)gridDim.x < p.n_items) len_next = item_len_raw(p, tab, item + gridDim.x);   // in flight during this item
      const int q = it.q0 + i * QT + row;
      // output row: (hi | lo) planes of DO columns each
      bf16* orow = p.out + ((long long)it.row0 + q) * (2 * DO) + it.h * HD + part * ND;
      if (it.dead || i >= it.n_act) {                 // tile of padded queries: defined zero output (no such rows when packed)
        if (!kRagged && q < p.T) {
#pragma unroll
          for (int c = 0; c < ND / 8; ++c) {
            *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(orow + DO + c * 8) = make_uint4(0, 0, 0, 0);
          }
        }
        continue;
      }
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < it.n_kt; ++j) {
        mbar_wait(&s_full[i], s_cnt & 1);             // S_i(j) ready; P_i(j-1) V retired: the P columns and O_i are ours
        ++s_cnt;
        tc_fence_after();
        uint32_t s[NC];
        tmem_ld64(s_addr + part * NC, s);
        tmem_ld_wait();
        const int valid = it.len - j * KT - part * NC;      // valid keys among this thread's columns (may be <= 0)
        if (valid < NC) {
#pragma unroll
          for (int c = 0; c < NC; ++c) if (c >= valid) s[c] = 0xff800000u;   // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < NC; c += 4) {
          mx[0] = fmaxf(mx[0], __uint_as_float(s[c])); mx[1] = fmaxf(mx[1], __uint_as_float(s[c + 1]));
          mx[2] = fmaxf(mx[2], __uint_as_float(s[c + 2])); mx[3] = fmaxf(mx[3], __uint_as_float(s[c + 3]));
        }
        float mxl = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * kLog2e;
        {   // combine the two parts' maxima; behind this barrier both parts hold their S columns in registers, so either
            // may overwrite the other's columns with its P planes
          const uint32_t slot = xch + (x_cnt & 1) * (SPLIT * QT * 4);
          ++x_cnt;
          sts_f32(slot + part * QT * 4, mxl);
          named_bar_sync(pair_bar, 32 * SPLIT);
          mxl = fmaxf(mxl, lds_f32(slot + (part ^ 1) * QT * 4));
        }
        float scale = 1.0f;
        const bool grow = mxl > m_used + kRescaleThreshold;     // always true for j == 0 (m_used = -inf)
        if (grow) { scale = ex2(m_used - mxl); m_used = mxl; }   // j == 0: scale = 0, l = 0, O not yet written
        if (j > 0 && __any_sync(0xffffffffu, grow)) {            // same rows, hence the same votes, in both parts' warps
          uint32_t o[ND];
          tmem_ld_n(o_addr, o);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < ND; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * scale);
          tmem_st_n(o_addr, o);
        }
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        // this thread's 64 keys = 32 packed columns per plane: P_h at [32 part, 32 part + 32), P_l 64 columns further
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const int cc = half * 32 + c;
            const float p0 = ex2(fmaf(__uint_as_float(s[cc]), kLog2e, -m_used));
            const float p1 = ex2(fmaf(__uint_as_float(s[cc + 1]), kLog2e, -m_used));
            const float p2 = ex2(fmaf(__uint_as_float(s[cc + 2]), kLog2e, -m_used));
            const float p3 = ex2(fmaf(__uint_as_float(s[cc + 3]), kLog2e, -m_used));
            sum[0] += p0; sum[1] += p1; sum[2] += p2; sum[3] += p3;
            split2(p0, p1, hi[c >> 1], lo[c >> 1]);
            split2(p2, p3, hi[(c >> 1) + 1], lo[(c >> 1) + 1]);
          }
          tmem_st16(s_addr + part * (NC / 2) + half * 16, hi);
          tmem_st16(s_addr + 64 + part * (NC / 2) + half * 16, lo);
        }
        l = fmaf(l, scale, (sum[0] + sum[1]) + (sum[2] + sum[3]));   // this part's share of the row sum
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[i]);
      }
      // ---- final: O_i / l -> (hi | lo) bf16 planes
      mbar_wait(&o_done[i], o_cnt & 1);
      ++o_cnt;
      tc_fence_after();
      uint32_t o[ND];
      tmem_ld_n(o_addr, o);
      tmem_ld_wait();
      {   // total row sum over the two parts
        const uint32_t slot = xch + (x_cnt & 1) * (SPLIT * QT * 4);
        ++x_cnt;
        sts_f32(slot + part * QT * 4, l);
        named_bar_sync(pair_bar, 32 * SPLIT);
        l += lds_f32(slot + (part ^ 1) * QT * 4);
      }
      const float inv = (q < it.len && l > 0.f) ? 1.0f / l : 0.f;     // padded query rows -> 0
      if (q < (kRagged ? it.len : p.T)) {
#pragma unroll
        for (int c = 0; c < ND / 8; ++c) {
          uint4 h4, l4;
          split2(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv, h4.x, l4.x);
          split2(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv, h4.y, l4.y);
          split2(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv, h4.z, l4.z);
          split2(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv, h4.w, l4.w);
          *reinterpret_cast<uint4*>(orow + c * 8) = h4;
          *reinterpret_cast<uint4*>(orow + DO + c * 8) = l4;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <typename TAB>
int attention_tc_x3_launch(const bf16* planes, bf16* out, const long long* lens, long long total_rows, int nb, int T, int H,
                           const TAB& tab, int num_sms, cudaStream_t s) {
  SWC_REQUIRE(((uintptr_t)planes & 15) == 0 && ((uintptr_t)out & 15) == 0, "attention_tc_x3: buffers must be 16-byte aligned");
  CUtensorMap tm;
  {
    cuuint64_t dims[2] = {(cuuint64_t)(6 * H * HD), (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {(cuuint64_t)(6 * H * HD) * 2};
    cuuint32_t box[2] = {HD, 128};
    SWC_TRY(make_tmap(&tm, 1, planes, 2, dims, strides, box));
  }
  AttnParams p{};
  p.lens = lens; p.out = out; p.T = T; p.H = H; p.nb = nb;
  p.n_qp = ceil_div(T, 2 * QT);
  p.n_items = p.n_qp * H * nb;
  SWC_REQUIRE((long long)p.n_items * std::max(p.n_qp * H, 1) < (1ll << 32), "attention_tc_x3: too many work items (%d)", p.n_items);
  p.by_qp.set((uint32_t)p.n_qp); p.by_h.set((uint32_t)H); p.by_qph.set((uint32_t)(p.n_qp * H));
  const int grid = std::min(p.n_items, num_sms);
  auto kern = attention_tc_x3_kernel<TAB>;
  SWC_TRY(ensure_dynamic_smem((const void*)kern, kSmemBytes));
  ProfScope ps(KC_ATTN, s);
  kern<<<grid, kThreads, kSmemBytes, s>>>(tm, p, tab);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int attention_tc_x3(const bf16* planes, bf16* out_planes, const long long* lens, int nb, int T, int H, int num_sms, cudaStream_t s) {
  return attention_tc_x3_launch(planes, out_planes, lens, (long long)nb * T, nb, T, H, NoTable{}, num_sms, s);
}

int attention_tc_x3_ragged(const bf16* planes, bf16* out_planes, const RaggedTable& tab, int H, int num_sms, cudaStream_t s) {
  SWC_REQUIRE(tab.nb > 0 && tab.nb <= kMaxRagged && tab.total > 0 && tab.t_max > 0, "attention_tc_x3_ragged: bad table");
  return attention_tc_x3_launch(planes, out_planes, nullptr, (long long)tab.total, tab.nb, tab.t_max, H, tab, num_sms, s);
}

}  // namespace swc

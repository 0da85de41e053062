// bf16 flash attention on tcgen05 (reference audiocodec/nn/modules.py:145-187): non-causal, head_dim 64, keys >=
// lens[b] masked, q pre-scaled by the packed q_proj.  Persistent, warp-specialised, two 128-query tiles per work item:
//
//   warp 0      TMA producer: Q tiles (double-buffered per item) and a 3-stage ring of K|V tiles (128 keys x 64 x bf16 each)
//   warp 1, 2   MMA issuer of query tile 0 / 1 (one lane each; warp 1 also owns the TMEM allocation)
//                 S_i = Q_i K_j^T   UMMA M128 N128 K16 x4, both operands K-major from shared memory -> TMEM (fp32)
//                 O_i += P_i V_j    UMMA M128 N64  K16 x8, A = P_i (bf16) from TMEM, B = V_j MN-major from shared memory
//               one issuer per tile: each blocks only on its own tile's barriers (a single issuer either serialises the
//               two warpgroups or pays ~150 cycles per polled barrier), and S runs one key block ahead of P V.
//   warp 3      idle (keeps the producer / issuer warps a whole warpgroup for setmaxnreg)
//   warps 4-11  softmax warps of query tile 0, warps 12-19 of query tile 1: two threads share a query row (= one TMEM lane),
//               each owns 64 of the block's 128 score columns and 32 of the 64 output dims; the halves of a row exchange
//               their maximum and their sum through shared memory behind a 64-thread named barrier of their own.
//   Registers: launched at 96 per thread; warpgroup 0 shrinks to 48 and the four softmax warpgroups grow to 104 (setmaxnreg),
//   which keeps 64 scores + 32 packed probabilities + the loop state of a softmax thread out of local memory.
//
// The two tiles ping-pong: while one warpgroup exponentiates S_i(j) the tensor core computes S_{1-i} / P V of the other.
// Per tile and key block a softmax thread reads its 128 scores from TMEM, takes the row maximum, and only when the
// maximum grew by more than 2^8 rescales its O row in TMEM (lazy rescale: the stale maximum is used otherwise, so
// probabilities stay <= 256 and the final 1/l normalisation is exact); it then writes P as packed bf16 back to TMEM.
// The normalised output tile is staged in shared memory and written as whole 128-byte rows (a thread-per-row store
// touches 32 different lines per instruction).
// TMEM columns: S0 0-127 | S1 128-255 | P0 256-319 | P1 320-383 | O0 384-447 | O1 448-511.
#include <algorithm>
#include <cstdlib>

#include <type_traits>

#include "attention_common.cuh"

namespace swc {

namespace {

using namespace ptx;
using namespace attn;

constexpr int STAGES = 3;
constexpr int kQOff = 0, kKVOff = 4 * kTileBytes;      // Q: [tile][buffer]
constexpr int kBarOff = kKVOff + STAGES * 2 * kTileBytes;
constexpr int kNumBars = 4 + 4 + 2 * STAGES + 2 + 2 + 2 + 2 + 2;
constexpr int kXchOff = kBarOff + kNumBars * 8 + 16;     // float [tile][buffer][part][row] exchange slots (SPLIT = 2)
constexpr int kOutOff = (kXchOff + 2 * 2 * 2 * QT * 4 + 127) / 128 * 128;   // bf16 output tiles [tile][128 rows][128 B], XOR-swizzled
constexpr int kSmemBytes = kOutOff + 2 * kTileBytes + 1024;
constexpr uint32_t kColS = 0, kColP = 256, kColO = 384;

template <int SPLIT, typename TAB>
__global__ void __launch_bounds__(128 + 256 * SPLIT, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p, const __grid_constant__ TAB tab) {
  constexpr bool kRagged = std::is_same<TAB, RaggedTable>::value;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* q_full = bars;               // [tile][buffer]
  uint64_t* q_empty = q_full + 4;        // [tile][buffer]
  uint64_t* kv_full = q_empty + 4;       // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES; // [STAGES]
  uint64_t* s_full = kv_empty + STAGES;  // [2] S_i written by the tensor core
  uint64_t* s_free = s_full + 2;         // [2] S_i copied to registers by its warpgroup (4 warps)
  uint64_t* p_full = s_free + 2;         // [2] P_i (and any O_i rescale) written by the warpgroup
  uint64_t* o_done = p_full + 2;         // [2] P_i V accumulated
  uint64_t* o_free = o_done + 2;         // [2] final O_i copied to registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4 * SPLIT);
      mbar_init(&p_full[i], 4 * SPLIT); mbar_init(&o_done[i], 1); mbar_init(&o_free[i], 4 * SPLIT);
    }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); }   // both tiles' issuers release a slot
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  const int D3 = 3 * p.H * HD;

  if (warp < 4) setmaxnreg_dec<kRegsIssue>();
  else setmaxnreg_inc<SPLIT == 2 ? kRegsSoftmax : kRegsSoftmax1>();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp in uniform control flow; one elected lane issues: see tc_ptx.cuh::elect_one)
    {
      uint32_t q_cnt[2] = {0, 0};
      uint32_t kv_it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode_item(p, tab, item);
        if (it.dead) continue;
        const int row0 = it.row0;
        for (int i = 0; i < it.n_act; ++i) {
          const uint32_t qb = i * 2 + (q_cnt[i] & 1);
          mbar_wait(&q_empty[qb], ((q_cnt[i] >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&q_full[qb], kTileBytes);
            tma_load_2d<1>(&tmQKV, smem_u32(&q_full[qb]), smem + kQOff + qb * kTileBytes, it.h * HD, row0 + it.q0 + i * QT);
          }
          __syncwarp();
          ++q_cnt[i];
        }
        for (int j = 0; j < it.n_kt; ++j, ++kv_it) {
          const uint32_t st = kv_it % STAGES, ph = (kv_it / STAGES) & 1;
          mbar_wait(&kv_empty[st], ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
            uint8_t* dst = smem + kKVOff + st * 2 * kTileBytes;
            tma_load_2d<1>(&tmQKV, smem_u32(&kv_full[st]), dst, (p.H + it.h) * HD, row0 + j * KT);
            tma_load_2d<1>(&tmQKV, smem_u32(&kv_full[st]), dst + kTileBytes, (2 * p.H + it.h) * HD, row0 + j * KT);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp <= 2) {
    // ------------------------------------------------------------------ MMA issuer of query tile i
    {
      const int i = warp - 1;
      constexpr uint32_t idesc_s = make_idesc(QT, KT);          // Q K^T: both K-major
      constexpr uint32_t idesc_o = make_idesc(QT, HD, 1);       // P V: B = V is MN-major ([key][dim], dim contiguous)
      // cursor over the (item, key block) pairs of this CTA, in the producer's ring order
      struct Cursor {
        int item, j, n_kt;
        bool active;          // tile i takes part in `item`
        uint32_t g;           // ring index of (item, j)
        uint32_t n_items;     // items with tile i active that lie behind the cursor
      };
      auto load_item = [&](Cursor& c) {            // position on the first non-dead item >= c.item
        for (; c.item < p.n_items; c.item += gridDim.x) {
          const Item it = decode_item(p, tab, c.item);
          if (it.dead) continue;
          c.n_kt = it.n_kt; c.active = i < it.n_act; c.j = 0;
          return;
        }
      };
      auto advance = [&](Cursor& c) {
        ++c.g;
        if (++c.j == c.n_kt) {
          if (c.active) ++c.n_items;
          c.item += gridDim.x;
          load_item(c);
        }
      };
      Cursor cs{blockIdx.x, 0, 0, false, 0, 0}, cp{blockIdx.x, 0, 0, false, 0, 0};
      load_item(cs);
      load_item(cp);
      uint32_t n_s = 0, n_p = 0;
      auto skip_inactive = [&](Cursor& c) {        // the S cursor only visits items in which this tile is active
        while (c.item < p.n_items && !c.active) { c.g += c.n_kt; c.item += gridDim.x; load_item(c); }
      };
      auto issue_s = [&]() {
        const uint32_t slot = cs.g % STAGES, qb = i * 2 + (cs.n_items & 1);
        if (cs.j == 0) mbar_wait(&q_full[qb], (cs.n_items >> 1) & 1);
        mbar_wait(&kv_full[slot], (cs.g / STAGES) & 1);
        mbar_wait(&s_free[i], (n_s & 1) ^ 1);                   // the previous S_i has been copied out by the warpgroup
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(smem + kQOff + qb * kTileBytes));
        const uint64_t bdesc = make_smem_desc(smem_u32(smem + kKVOff + slot * 2 * kTileBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16<1>(tmem_base + kColS + i * 128, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
          umma_commit<1>(&s_full[i]);
          if (cs.j + 1 == cs.n_kt) umma_commit<1>(&q_empty[qb]);  // last use of this Q buffer
        }
        __syncwarp();
        ++n_s;
        advance(cs);
        skip_inactive(cs);
      };
      skip_inactive(cs);
      while (cp.item < p.n_items) {
        const uint32_t slot = cp.g % STAGES;
        // S runs at most one ring entry ahead of the P V cursor.  (Never further: the producer can only fill a slot
        // once BOTH issuers have released the entry STAGES before it, and this issuer releases in cursor order.)
        while (cs.item < p.n_items && cs.g <= cp.g + 1) issue_s();
        if (!cp.active) {
          // this tile sits the item out but still owes the ring its release: in ring order, once the slot has landed
          mbar_wait(&kv_full[slot], (cp.g / STAGES) & 1);
          if (elect_one()) mbar_arrive(&kv_empty[slot]);
          __syncwarp();
          advance(cp);
          continue;
        }
        mbar_wait(&p_full[i], n_p & 1);
        if (cp.j == 0) mbar_wait(&o_free[i], (cp.n_items & 1) ^ 1);   // the previous item's O_i has been read out
        tc_fence_after();
        const uint64_t vdesc = make_smem_desc(smem_u32(smem + kKVOff + slot * 2 * kTileBytes + kTileBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) {
            // A: 16 keys = 8 packed columns of P_i; B: 16 key rows of V = 2 swizzle atoms of 8 rows x 128 B (+128 in >>4 units)
            umma_bf16_ts(tmem_base + kColO + i * 64, tmem_base + kColP + i * 64 + k * 8, vdesc + (uint64_t)(k * 128), idesc_o,
                         (cp.j | k) != 0);
          }
          umma_commit<1>(&o_done[i]);
          umma_commit<1>(&kv_empty[slot]);
        }
        __syncwarp();
        ++n_p;
        advance(cp);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax warps
    // SPLIT threads share one query row: thread `part` owns key columns [part*NC, part*NC + NC) of S / P and output
    // dims [part*ND, part*ND + ND) of O.  With SPLIT = 2 there are 16 softmax warps (4 per scheduler instead of 2)
    // to hide the MUFU / TMEM latencies; the row maximum and the final row sum are combined through shared memory.
    constexpr int NC = KT / SPLIT, ND = HD / SPLIT;
    const int sw = warp - 4;                          // softmax warp index
    const int i = sw / (4 * SPLIT);                   // query tile
    const int part = (sw >> 2) % SPLIT;               // column part of the row
    const int lg = warp & 3;                          // TMEM lane group this warp may access
    // the SPLIT warps that share a row group exchange maxima / sums through a barrier of their own (they sit on the same
    // scheduler); only the output staging needs the whole tile (barrier 9 + i)
    const int pair_bar = 1 + i * 4 + lg;
    const int row = lg * 32 + lane;                   // query row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    const int DO = p.H * HD;
    const uint32_t xch = smem_u32(smem + kXchOff) + (uint32_t)(i * (2 * SPLIT * QT) + row) * 4u;   // [2 buffers][SPLIT][128 rows]
    uint32_t s_cnt = 0, o_cnt = 0, x_cnt = 0;
    int tr = 0;
    const bool tracer = p.trace != nullptr && lg == 3 && part == 0 && lane == 0;   // first softmax warp of each tile
    const int tslot = i;
    uint8_t* ostage = smem + kOutOff + i * kTileBytes;
    long long len_next = blockIdx.x < p.n_items ? item_len_raw(p, tab, blockIdx.x) : 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item(p, tab, item, len_next);
      if (item + (int)gridDim.x < p.n_items) len_next = item_len_raw(p, tab, item + gridDim.x);   // in flight during this item
      const int q = it.q0 + i * QT + row;
      bf16* orow = p.out + ((long long)it.row0 + q) * DO + it.h * HD + part * ND;
      if (it.dead || i >= it.n_act) {                 // tile of padded queries: defined zero output (no such rows when packed)
        if (!kRagged && q < p.T) {
#pragma unroll
          for (int c = 0; c < ND / 8; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      float m_used = -INFINITY, l = 0.f;
      if (tracer) trace_stamp(p, tslot, tr);                 // A: item decoded
      const uint32_t sa = lane_addr + kColS + i * 128 + part * NC;
      constexpr int NV = NC / 32;                          // 32-register score vectors per thread (2 or 4)
      for (int j = 0; j < it.n_kt; ++j) {
        mbar_wait(&s_full[i], s_cnt & 1);
        if (tracer) trace_stamp(p, tslot, tr);               // B: S ready
        ++s_cnt;
        tc_fence_after();
        uint32_t sv[NV][32];
#pragma unroll
        for (int v = 0; v < NV; ++v) tmem_ld32(sa + 32 * v, sv[v]);
        tmem_ld_wait();
        if (tracer && p.fine) trace_stamp(p, tslot, tr);               // B1: S in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[i]);
        // a barrier poll costs ~200 cycles of latency even when its phase is long complete: poll now, consume behind the exponentials
        const bool pv_early = j > 0 && mbar_test(&o_done[i], o_cnt & 1);
        const int valid = it.len - j * KT - part * NC;      // valid keys among this thread's columns (may be <= 0)
        if (valid < NC) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (v * 32 + c >= valid) sv[v][c] = 0xff800000u;   // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int v = 0; v < NV; v += 2)
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            mx[0] = fmaxf(mx[0], fmaxf(__uint_as_float(sv[v][c]), __uint_as_float(sv[v + 1][c])));
            mx[1] = fmaxf(mx[1], fmaxf(__uint_as_float(sv[v][c + 1]), __uint_as_float(sv[v + 1][c + 1])));
            mx[2] = fmaxf(mx[2], fmaxf(__uint_as_float(sv[v][c + 2]), __uint_as_float(sv[v + 1][c + 2])));
            mx[3] = fmaxf(mx[3], fmaxf(__uint_as_float(sv[v][c + 3]), __uint_as_float(sv[v + 1][c + 3])));
          }
        float mxl = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * kLog2e;
        if constexpr (SPLIT > 1) {                    // combine the parts' maxima (double-buffered exchange slots)
          const uint32_t slot = xch + (x_cnt & 1) * (SPLIT * QT * 4);
          ++x_cnt;
          sts_f32(slot + part * QT * 4, mxl);
          named_bar_sync(pair_bar, 32 * SPLIT);
          mxl = fmaxf(mxl, lds_f32(slot + (part ^ 1) * QT * 4));
        }
        if (tracer && p.fine) trace_stamp(p, tslot, tr);               // B2: row maximum known
        float scale = 1.0f;
        const bool grow = mxl > m_used + kRescaleThreshold;     // always true for j == 0 (m_used = -inf)
        if (grow) { scale = ex2(m_used - mxl); m_used = mxl; }   // j == 0: scale = 0, l = 0, O not yet written
        uint32_t pk[NV][16];
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float p0 = ex2(fmaf(__uint_as_float(sv[v][c]), kLog2e, -m_used));
            const float p1 = ex2(fmaf(__uint_as_float(sv[v][c + 1]), kLog2e, -m_used));
            const float p2 = ex2(fmaf(__uint_as_float(sv[v][c + 2]), kLog2e, -m_used));
            const float p3 = ex2(fmaf(__uint_as_float(sv[v][c + 3]), kLog2e, -m_used));
            sum[0] += p0; sum[1] += p1; sum[2] += p2; sum[3] += p3;
            pk[v][c >> 1] = pack2(p0, p1);
            pk[v][(c >> 1) + 1] = pack2(p2, p3);
          }
        l = fmaf(l, scale, (sum[0] + sum[1]) + (sum[2] + sum[3]));   // this part's share of the row sum
        if (tracer && p.fine) trace_stamp(p, tslot, tr);               // B3: exponentials done
        if (j > 0) {
          if (!pv_early) mbar_wait(&o_done[i], o_cnt & 1);   // P_i(j-1) V accumulated: P_i is free, O_i is stable
          ++o_cnt;
          tc_fence_after();
          if (__any_sync(0xffffffffu, grow)) {          // same rows, hence the same votes, in every part's warp
            const uint32_t oa = lane_addr + kColO + i * 64 + part * ND;
#pragma unroll
            for (int h2 = 0; h2 < ND / 32; ++h2) {
              uint32_t o[32];
              tmem_ld32(oa + 32 * h2, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * scale);
              tmem_st32(oa + 32 * h2, o);
            }
          }
        }
        if (tracer && p.fine) trace_stamp(p, tslot, tr);               // B4: previous P V done (P slot free)
        {
          const uint32_t pa = lane_addr + kColP + i * 64 + part * (NC / 2);
#pragma unroll
          for (int v = 0; v < NV; ++v) tmem_st16(pa + 16 * v, pk[v]);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[i]);
        if (tracer) trace_stamp(p, tslot, tr);               // C: P handed over
      }
      // ---- final: O_i / l -> bf16 rows
      mbar_wait(&o_done[i], o_cnt & 1);
      if (tracer) trace_stamp(p, tslot, tr);                 // D: last P V done
      ++o_cnt;
      tc_fence_after();
      uint32_t o[ND];
      if constexpr (ND == 32) {
        tmem_ld32(lane_addr + kColO + i * 64 + part * ND, o);
      } else {
        tmem_ld64(lane_addr + kColO + i * 64, o);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[i]);
      if constexpr (SPLIT > 1) {                        // total row sum over the parts
        const uint32_t slot = xch + (x_cnt & 1) * (SPLIT * QT * 4);
        ++x_cnt;
        sts_f32(slot + part * QT * 4, l);
        named_bar_sync(pair_bar, 32 * SPLIT);
        l += lds_f32(slot + (part ^ 1) * QT * 4);
      }
      const float inv = (q < it.len && l > 0.f) ? 1.0f / l : 0.f;     // padded query rows -> 0
      {
        // this thread's ND dims of row `row` -> staging tile (16-byte chunk c of row r lives at chunk c ^ (r & 7))
        const uint32_t srow = smem_u32(ostage) + (uint32_t)row * 128u;
#pragma unroll
        for (int c = 0; c < ND / 8; ++c) {
          const uint32_t chunk = (uint32_t)(part * (ND / 8) + c);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((chunk ^ (uint32_t)(row & 7)) << 4)),
                       "r"(pack2(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv)),
                       "r"(pack2(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv)),
                       "r"(pack2(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv)),
                       "r"(pack2(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv))
                       : "memory");
        }
        named_bar_sync(9 + i, 128 * SPLIT);
        // whole 128-byte rows: each of the tile's 4*SPLIT warps writes 128 / (4*SPLIT) rows, 4 rows per instruction
        constexpr int kRowsPerWarp = QT / (4 * SPLIT);
        const int wt = sw % (4 * SPLIT);                 // warp index inside the tile
        bf16* obase = p.out + ((long long)it.row0 + it.q0 + i * QT) * DO + it.h * HD;
        const int row_limit = (kRagged ? it.len : p.T) - (it.q0 + i * QT);     // rows of this tile that exist
#pragma unroll
        for (int r4 = 0; r4 < kRowsPerWarp / 4; ++r4) {
          const int r = wt * kRowsPerWarp + r4 * 4 + (lane >> 3), c = lane & 7;
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(smem_u32(ostage) + (uint32_t)r * 128u + (((uint32_t)c ^ (uint32_t)(r & 7)) << 4)));
          if (r < row_limit) *reinterpret_cast<uint4*>(obase + (long long)r * DO + c * 8) = v;
        }
      }
      if (tracer) trace_stamp(p, tslot, tr);                 // E: item stored
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
  (void)D3;
}

}  // namespace

long long* g_attn_trace = nullptr;   // set by swc_debug_attn_trace
int g_attn_trace_fine = 0;

namespace {
template <typename TAB>
int attention_tc_launch(const bf16* qkv, bf16* out, const long long* lens, long long total_rows, int nb, int T, int H,
                        const TAB& tab, int num_sms, cudaStream_t s) {
  SWC_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "attention_tc: buffers must be 16-byte aligned");
  CUtensorMap tm;
  {
    cuuint64_t dims[2] = {(cuuint64_t)(3 * H * HD), (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {(cuuint64_t)(3 * H * HD) * 2};
    cuuint32_t box[2] = {HD, 128};
    SWC_TRY(make_tmap(&tm, 1, qkv, 2, dims, strides, box));
  }
  AttnParams p{};
  p.lens = lens; p.out = out; p.T = T; p.H = H; p.nb = nb;
  p.n_qp = ceil_div(T, 2 * QT);
  p.n_items = p.n_qp * H * nb;
  SWC_REQUIRE((long long)p.n_items * std::max(p.n_qp * H, 1) < (1ll << 32), "attention_tc: too many work items (%d)", p.n_items);
  p.by_qp.set((uint32_t)p.n_qp); p.by_h.set((uint32_t)H); p.by_qph.set((uint32_t)(p.n_qp * H));
  p.trace = g_attn_trace;
  p.fine = g_attn_trace_fine;
  const int grid = std::min(p.n_items, num_sms);
  ProfScope ps(KC_ATTN, s);
  auto go = [&](auto kern, int threads) -> int {
    SWC_TRY(ensure_dynamic_smem((const void*)kern, kSmemBytes));
    kern<<<grid, threads, kSmemBytes, s>>>(tm, p, tab);
    return 0;
  };
  // two threads per query row.  (One thread per row, SPLIT = 1 with 208 registers per softmax thread, is correct and 4 %
  // slower: a single warp per scheduler and tile reaches only ~62 % of the MUFU rate, tools/micro/mufu_rate.cu.)
  const int rc = go(attention_tc_kernel<2, TAB>, kThreads);      // two threads per query row (the one-thread-per-row form is gone)
  if (rc) return rc;
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

int attention_tc(const bf16* qkv, bf16* out, const long long* lens, int nb, int T, int H, int num_sms, cudaStream_t s) {
  return attention_tc_launch(qkv, out, lens, (long long)nb * T, nb, T, H, NoTable{}, num_sms, s);
}

int attention_tc_ragged(const bf16* qkv, bf16* out, const RaggedTable& tab, int H, int num_sms, cudaStream_t s) {
  SWC_REQUIRE(tab.nb > 0 && tab.nb <= kMaxRagged && tab.total > 0 && tab.t_max > 0, "attention_tc_ragged: bad table");
  return attention_tc_launch(qkv, out, nullptr, (long long)tab.total, tab.nb, tab.t_max, H, tab, num_sms, s);
}

}  // namespace swc

// ---- debug: per-phase clock stamps of CTA 0 (tools/attn_trace.py) ----------------------------------------------------
extern "C" int swc_debug_attn_trace(int enable, long long* host_out, int n) {
  using namespace swc;
  if (enable) {
    if (!g_attn_trace && cudaMalloc(&g_attn_trace, 3 * 4096 * sizeof(long long)) != cudaSuccess) return -1;
    cudaMemset(g_attn_trace, 0, 3 * 4096 * sizeof(long long));
    g_attn_trace_fine = (enable & 2) ? 1 : 0;
    return 0;
  }
  if (!g_attn_trace) return -1;
  cudaDeviceSynchronize();
  if (host_out) cudaMemcpy(host_out, g_attn_trace, sizeof(long long) * (size_t)(n < 3 * 4096 ? n : 3 * 4096), cudaMemcpyDeviceToHost);
  cudaFree(g_attn_trace);
  g_attn_trace = nullptr;
  return 0;
}

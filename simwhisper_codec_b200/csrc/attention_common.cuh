// Pieces shared by the two tcgen05 flash-attention kernels (attention_tc.cu: bf16; attention_tc_x3.cu: three-product
// split-bf16, fp32-class): the persistent work-item decode (batch, head, pair of 128-query tiles), softmax helpers.
#pragma once
#include <type_traits>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace swc {
namespace attn {

using namespace ptx;

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int kTileBytes = 128 * 64 * 2;                 // 16 KB: one Q, K or V tile (128 rows x 64 bf16)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;                // log2 units
// Thread layout of both kernels: warpgroup 0 = warp 0 TMA producer, warps 1 / 2 MMA issuers of query tile 0 / 1, warp 3 idle;
// warpgroups 1-4 = the 16 softmax warps.  The kernels are launched at 96 registers per thread (640 x 96 = 60 K of the 64 K
// register file); warpgroup 0 shrinks to kRegsIssue and hands its registers to the softmax warpgroups, which grow to
// kRegsSoftmax (128 x 48 + 512 x 104 = 58 K): the softmax threads hold 64 scores, 32 packed probabilities and the loop state.
constexpr int kThreads = 128 + 512;
constexpr int kRegsIssue = 48, kRegsSoftmax = 104;
constexpr int kRegsSoftmax1 = 208;      // one thread per query row (experiment): 384 threads launched at 168 registers
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// n / d by one multiply-high (exact while n * d < 2^32): the per-item decode runs on every role's critical path
struct FastDiv {
  uint32_t mul, d;
  __host__ void set(uint32_t dd) { d = dd; mul = (uint32_t)(((1ull << 32) + dd - 1) / dd); }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : __umulhi(n, mul); }
};
struct AttnParams {
  const long long* lens;   // device lengths (padded layout) or null
  bf16* out;
  int T, H, nb, n_qp, n_items;
  FastDiv by_qp, by_h, by_qph;
  long long* trace;        // optional per-phase clock64 stamps of CTA 0 (tools/attn_trace.py); null in normal runs
  int fine;                // trace: four more stamps per key block (perturbs the kernel)
};
// RAGGED: rows are packed (item b = rows [off[b], off[b] + len[b])), lengths come with the launch (no global load per
// item), and rows >= len[b] do not exist: nothing is written for them.
struct NoTable {};

struct Item {
  int b, h, q0, len, n_act, n_kt, row0;
  bool dead;
};
// raw length of the item's sequence: the only memory access of the decode, split off so that it can be issued one item
// ahead.  Nothing may depend on the loaded value before the next item starts (warps issue in order: a dependent
// instruction right behind the load would stall for the whole L2 latency), so the clamp to [0, T] lives in decode_item.
template <typename TAB>
__device__ __forceinline__ long long item_len_raw(const AttnParams& p, const TAB& tab, int item) {
  const int b = (int)p.by_qph.div((uint32_t)item);
  if constexpr (std::is_same<TAB, RaggedTable>::value) {
    return tab.len[b];
  } else {
    return p.lens ? p.lens[b] : (long long)p.T;
  }
}
template <typename TAB>
__device__ __forceinline__ Item decode_item(const AttnParams& p, const TAB& tab, int item, long long len_raw) {
  Item it;
  const int r = (int)p.by_qp.div((uint32_t)item);
  const int qp = item - r * p.n_qp;
  it.b = (int)p.by_h.div((uint32_t)r);
  it.h = r - it.b * p.H;
  it.q0 = qp * 2 * QT;
  if constexpr (std::is_same<TAB, RaggedTable>::value) {
    it.len = (int)len_raw;
    it.row0 = tab.off[it.b];
  } else {
    it.len = (int)(len_raw > p.T ? p.T : (len_raw < 0 ? 0 : len_raw));
    it.row0 = it.b * p.T;
  }
  it.len = (int)uniform_u32((uint32_t)it.len);      // same address in every lane: tell the compiler it is warp-uniform
  it.row0 = (int)uniform_u32((uint32_t)it.row0);
  it.dead = it.q0 >= it.len;
  it.n_act = (it.q0 + QT < it.len) ? 2 : 1;
  it.n_kt = (it.len + KT - 1) / KT;
  return it;
}
template <typename TAB>
__device__ __forceinline__ Item decode_item(const AttnParams& p, const TAB& tab, int item) {
  return decode_item(p, tab, item, item_len_raw(p, tab, item));
}

__device__ __forceinline__ void trace_stamp(const AttnParams& p, int slot, int& idx) {
  // slots: 0 / 1 first softmax warp of query tile 0 / 1; 4096 stamps each
  if (p.trace && blockIdx.x == 0 && idx < 4096) p.trace[slot * 4096 + idx++] = clock64();
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// (no pointer casts into sub-arrays: they force the register array into local memory)
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st32(taddr, r); }

}  // namespace attn
}  // namespace swc

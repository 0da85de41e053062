// Shared device/host helpers for the SimWhisper-Codec sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

namespace swc {

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define SWC_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      swc::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)
#define SWC_REQUIRE(cond, ...)                                                                 \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      swc::set_error(__VA_ARGS__);                                                             \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)
#define SWC_TRY(expr)                                                                          \
  do {                                                                                         \
    int _r = (expr);                                                                           \
    if (_r != 0) return _r;                                                                    \
  } while (0)

// opt a kernel into `bytes` of dynamic shared memory on the current device (once per device and function; gemm_tc2.cu)
int ensure_dynamic_smem(const void* func, int bytes);

// makes `device` current for the lifetime of the object and restores the caller's device afterwards: every C entry point
// launches on the model's device whatever the process's current device is (torch.cuda.set_device elsewhere, threads, ...)
struct DeviceGuard {
  int prev = -1, want = -1;
  bool good = true;
  explicit DeviceGuard(int device) : want(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { good = false; return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) good = false;
  }
  ~DeviceGuard() { if (good && prev >= 0 && prev != want) cudaSetDevice(prev); }
  bool ok() const { return good; }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// ---- per-kernel-class launch counters and (optional) CUDA-event timers ------------------------------
enum KClass { KC_GEMM_TC = 0, KC_GEMM_SIMT, KC_ATTN, KC_LAYERNORM, KC_DWCONV_LN, KC_SNAKE, KC_MISC, KC_COUNT };
void prof_begin(int cls, cudaStream_t s);
void prof_end(int cls, cudaStream_t s);
struct ProfScope {
  int cls; cudaStream_t s;
  ProfScope(int c, cudaStream_t st) : cls(c), s(st) { prof_begin(cls, s); }
  ~ProfScope() { prof_end(cls, s); }
};

// ---- element conversion ------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// load 8 consecutive elements (16-byte aligned for bf16, 32-byte for float) as floats
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// bf16x3 mode: values leave a kernel as two bf16 planes, x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (2^-17 relative):
// hi at p[0..n), lo at p[plane_stride + 0..n).  Same arithmetic as split_bf16_planes (kernels.cu), so a producer that
// writes planes directly is bit-identical to writing fp32 and splitting afterwards.
struct bf16_planes {};      // output-type tag of the row-wise kernels: rows of (hi | lo) planes, C columns each
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void store8_planes(bf16* p, long long plane_stride, const float (&v)[8]) {
  uint4 h, l;
  split_pair(v[0], v[1], h.x, l.x); split_pair(v[2], v[3], h.y, l.y);
  split_pair(v[4], v[5], h.z, l.z); split_pair(v[6], v[7], h.w, l.w);
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + plane_stride) = l;
}
__device__ __forceinline__ void store4_planes(bf16* p, long long plane_stride, float4 v) {
  uint2 h, l;
  split_pair(v.x, v.y, h.x, l.x); split_pair(v.z, v.w, h.y, l.y);
  *reinterpret_cast<uint2*>(p) = h;
  *reinterpret_cast<uint2*>(p + plane_stride) = l;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact (erf) GELU, as torch.nn.functional.gelu default
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// float atomic max valid for any sign (target initialised to -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace swc

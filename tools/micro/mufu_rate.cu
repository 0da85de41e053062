// Micro-benchmark: sustained MUFU.EX2 rate per SM (alone, and mixed with the FFMA / FADD / F2FP of a softmax inner loop),
// for 4 / 8 / 16 warps per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -1.0f - 0.01f * (threadIdx.x + i);
  float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float x0 = a[i], x1 = a[i + 1];
      if (MODE >= 1) { x0 = fmaf(x0, 1.4426950408889634f, -0.25f); x1 = fmaf(x1, 1.4426950408889634f, -0.25f); }
      const float p0 = ex2(x0), p1 = ex2(x1);
      if (MODE >= 1) { s0 += p0; s1 += p1; }
      if (MODE >= 2) { __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1); acc ^= *reinterpret_cast<unsigned*>(&h); }
      a[i] = MODE == 0 ? p0 - 2.0f : a[i];
      a[i + 1] = MODE == 0 ? p1 - 2.0f : a[i + 1];
      if (MODE >= 1) { s2 += x0 * 1e-9f; s3 += x1 * 1e-9f; a[i] -= 1e-7f; a[i + 1] -= 1e-7f; }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] + s0 + s1 + s2 + s3 + (float)acc;
}

template <int MODE>
void run(int warps, const char* name) {
  const int blocks = 148, threads = warps * 32, iters = 2000;
  float* out; long long* cyc;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  const double ex = (double)iters * 8 * threads;
  printf("%-28s %2d warps/SM: %.2f ex2 per clk per SM  (%.1f cycles per warp-wide MUFU per scheduler)\n", name, warps, ex / avg, avg / (iters * 8.0 * warps / 4));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) run<0>(w, "MUFU.EX2 only (dependent)");
  for (int w : {4, 8, 16}) run<1>(w, "FFMA + EX2 + FADD");
  for (int w : {4, 8, 16}) run<2>(w, "FFMA + EX2 + FADD + F2FP");
  return 0;
}

// Throughput of the exponential flavours a softmax thread can choose from on sm_100a: MUFU ex2 on f32,
// packed f16x2 / bf16x2 ex2, and the FMA-pipe cubic.  Prints results per SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_ubench_ex2 tools/ubench_ex2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -127.f);
  const float xf = __fadd_rd(x, 12582912.f);
  const float f = x - (xf - 12582912.f);
  float p = fmaf(0.077119089663028717f, f, 0.227564394474029541f);
  p = fmaf(p, f, 0.695146143436431885f);
  p = fmaf(p, f, 1.f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}

template <int MODE>
__global__ void k(float *out, int iters, float seed) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed - 0.01f * (threadIdx.x + i); u[i] = 0xB800B800u + i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.5f;
      if (MODE == 1) u[i] = ex2h2(u[i]) ^ 0x80008000u;
      if (MODE == 2) u[i] = ex2b2(u[i]) ^ 0x80008000u;
      if (MODE == 3) a[i] = ex2_poly(a[i]) - 1.5f;
      if (MODE == 4) { a[i] = (i & 1) ? ex2_poly(a[i]) - 1.5f : ex2f(a[i]) - 1.5f; }
      if (MODE == 5) { a[i] = (i & 3) == 3 ? ex2_poly(a[i]) - 1.5f : ex2f(a[i]) - 1.5f; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, int per_op) {
  float *out; cudaMalloc(&out, 148 * 8 * 512 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 2, 512>>>(out, 100, -0.3f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 512>>>(out, iters, -0.3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double results = (double)iters * 8 * per_op * 1024;      // per SM (2 CTAs x 512 threads)
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s %8.3f ms  %6.2f results/clk/SM (at the nominal %d MHz)  err=%s\n", name, ms, results / cycles, clk / 1000,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

__global__ void acc(float *o) {   // accuracy of the cubic against exp2f
  float worst = 0.f;
  for (int i = 0; i < 100000; ++i) {
    const float x = -20.f * (i / 100000.f);
    const float r = exp2f(x), p = ex2_poly(x);
    worst = fmaxf(worst, fabsf(p - r) / r);
  }
  *o = worst;
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("cubic on the FMA pipe", 1);
  run<4>("1/2 cubic + 1/2 MUFU", 1);
  run<5>("1/4 cubic + 3/4 MUFU", 1);
  float *o; cudaMalloc(&o, 4); acc<<<1, 1>>>(o); float w; cudaMemcpy(&w, o, 4, cudaMemcpyDeviceToHost);
  printf("cubic max relative error on [-20, 0]: %.3g\n", w);
  return 0;
}

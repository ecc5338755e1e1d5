// Micro-benchmark: MUFU.EX2 issue rate per SM as a function of resident warps, and a degree-3 polynomial exp2 on the
// FMA pipe (tuning probe, not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {   // x <= 0
  x = fmaxf(x, -126.0f);
  const float xf = x + 12582912.0f;                 // round to nearest integer in the low mantissa bits
  const float n = xf - 12582912.0f;
  const float f = x - n;                            // [-0.5, 0.5]
  float p = fmaf(f, 0.0555041086f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
template <int MODE> __global__ void k(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed - i * 0.01f - threadIdx.x * 1e-4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2a(a[i]) - 1.5f;
      else if (MODE == 1) a[i] = ex2_poly(a[i]) - 1.5f;
      else { if (i & 3) a[i] = ex2a(a[i]) - 1.5f; else a[i] = ex2_poly(a[i]) - 1.5f; }   // 25 % offloaded
    }
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SMs %d clock %d kHz\n", sms, clk);
  const int iters = 4000;
  for (int mode = 0; mode < 3; ++mode)
    for (int threads = 128; threads <= 1024; threads *= 2) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<sms, threads>>>(out, iters, -0.3f);
        else if (mode == 1) k<1><<<sms, threads>>>(out, iters, -0.3f);
        else k<2><<<sms, threads>>>(out, iters, -0.3f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)sms * threads * iters * 16;
      printf("mode %d warps/SM %2d: %.3f ms  %.2f exp/clk/SM (at %.0f MHz nominal)\n", mode, threads / 32, ms,
             ops / sms / (ms * 1e-3) / (clk * 1e3), clk / 1e3);
    }
  return 0;
}

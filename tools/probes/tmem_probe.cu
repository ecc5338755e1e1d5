// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM for different shapes / warp counts (tuning probe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int N> struct Ld;
#define REGS8(b) "%" #b
template <> struct Ld<32> {
  static __device__ __forceinline__ void run(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(a) : "memory");
  }
};
template <> struct Ld<8> {
  static __device__ __forceinline__ void run(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(a) : "memory");
  }
};
// MODE 0: x32 loads, wait after every 4 (128 columns); MODE 1: x32, wait after each; MODE 2: x8 loads, wait after 16
template <int MODE> __global__ void k(float* out, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      uint32_t r[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) Ld<32>::run(base + c * 32, r + c * 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 128; ++i) acc ^= r[i];
    } else if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        Ld<32>::run(base + c * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    } else {
      uint32_t r[128];
#pragma unroll
      for (int c = 0; c < 16; ++c) Ld<8>::run(base + c * 8, r + c * 8);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 128; ++i) acc ^= r[i];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tptr) : "memory");
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int threads = 128; threads <= 512; threads *= 2) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<sms, threads>>>(out, iters);
        else if (mode == 1) k<1><<<sms, threads>>>(out, iters);
        else k<2><<<sms, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("mode %d warps %2d: %.3f ms -> %.1f B/clk/SM (%s)\n", mode, threads / 32, ms,
             (double)threads * iters * 128 * 4 / (ms * 1e-3) / (clk * 1e3), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}

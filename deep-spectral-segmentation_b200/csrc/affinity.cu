// Patch-affinity build: W = relu(F^ F^T) / max (+ lambda * colour counts)   (reference extract/extract.py:148,191-221)
//
// fp32 end to end: the eigenvectors of the graph Laplacian move by 2e-4..8e-4 when the features are rounded to
// bf16, so this stage does not use reduced-precision tensor-core operands.
#include "common.cuh"

namespace dss {

// ---- row normalisation: fn = f / max(||f||_2, 1e-12)  (F.normalize), and the per-image max of the diagonal of
// F^ F^T (== max of the whole matrix by Cauchy-Schwarz), accumulated with an integer atomicMax on the float bits.
__global__ void __launch_bounds__(256)
rownorm_kernel(const float* __restrict__ f, float* __restrict__ fn, unsigned int* __restrict__ img_max, int rows,
               int N, int d, int normalize) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = f + (long long)row * d;
  float* y = fn + (long long)row * d;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) ss = fmaf(x[k], x[k], ss);
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
  float s2 = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = normalize ? x[k] / denom : x[k];
    y[k] = v;
    s2 = fmaf(v, v, s2);
  }
  s2 = warp_sum(s2);
  if (lane == 0) atomicMax(img_max + row / N, __float_as_uint(fmaxf(s2, 0.f)));
}

// ---- W tile kernel: 64x64 tile per CTA, 16x16 threads, 4x4 micro-tile, K chunks of 16 through shared memory.
constexpr int AT = 64, AK = 16;

__global__ void __launch_bounds__(256)
affinity_kernel(const float* __restrict__ fn, const unsigned int* __restrict__ img_max,
                const uint8_t* __restrict__ counts, float lambda, float* __restrict__ Wm, int N, int d, int ldw,
                int threshold) {
  __shared__ float As[AK][AT + 4];
  __shared__ float Bs[AK][AT + 4];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * AT, c0 = blockIdx.x * AT;
  const float* F = fn + (long long)b * N * d;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row in tile, k offset
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d; k0 += AK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bb = a;
    if (r0 + lr < N && k0 + lk < d) a = *reinterpret_cast<const float4*>(F + (long long)(r0 + lr) * d + k0 + lk);
    if (c0 + lr < N && k0 + lk < d) bb = *reinterpret_cast<const float4*>(F + (long long)(c0 + lr) * d + k0 + lk);
    As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
    Bs[lk + 0][lr] = bb.x; Bs[lk + 1][lr] = bb.y; Bs[lk + 2][lr] = bb.z; Bs[lk + 3][lr] = bb.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < AK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float mx = __uint_as_float(img_max[b]);
  float* Wb = Wm + (long long)b * N * ldw;
  const uint8_t* Cb = counts ? counts + (long long)b * N * N : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= N) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      float w = acc[i][j];
      if (threshold) w = w > 0.f ? w : 0.f;        // W * (W > 0)
      w = w / mx;                                  // W / W.max()
      if (Cb && c < N) w += static_cast<float>(Cb[(long long)r * N + c]) * lambda;  // + W_color * lambda
      o[j] = (c < N) ? w : 0.f;                    // columns [N, ldw) are zero padding
    }
    const int c = c0 + tx * 4;
    if (c + 3 < ldw) {
      *reinterpret_cast<float4*>(Wb + (long long)r * ldw + c) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < ldw) Wb[(long long)r * ldw + c + j] = o[j];
    }
  }
}

}  // namespace dss

using namespace dss;

extern "C" size_t dss_affinity_workspace_bytes(int B, int N, int d) {
  if (B <= 0 || N <= 0 || d <= 0) return 0;
  return align_up((size_t)B * N * d * sizeof(float), 256) + align_up((size_t)B * sizeof(unsigned int), 256);
}

extern "C" int dss_affinity(const float* feats, int B, int N, int d, int flags, const uint8_t* color_counts,
                            float color_lambda, float* Wmat, int ldw, void* ws, size_t ws_bytes, dss_stream_t stream) {
  DSS_REQUIRE(feats && Wmat && ws, "affinity: null pointer");
  DSS_REQUIRE(B > 0 && N > 0 && d > 0, "affinity: empty problem B=%d N=%d d=%d", B, N, d);
  DSS_REQUIRE(d % 4 == 0, "affinity: feature dim must be a multiple of 4 (got %d)", d);
  DSS_REQUIRE(ldw >= N && ldw % 4 == 0, "affinity: ldw must be >= N and a multiple of 4 (N=%d ldw=%d)", N, ldw);
  DSS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "affinity: workspace must be 256-byte aligned");
  if (ws_bytes < dss_affinity_workspace_bytes(B, N, d)) {
    set_error("affinity: workspace too small (%zu < %zu)", ws_bytes, dss_affinity_workspace_bytes(B, N, d));
    return DSS_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* fn = reinterpret_cast<float*>(ws);
  unsigned int* img_max =
      reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(ws) + align_up((size_t)B * N * d * sizeof(float), 256));
  DSS_CHECK_CUDA(cudaMemsetAsync(img_max, 0, (size_t)B * sizeof(unsigned int), st));
  const int rows = B * N;
  {
    LaunchScope scope(st, KC_ROWNORM);
    rownorm_kernel<<<cdiv(rows, 8), 256, 0, st>>>(feats, fn, img_max, rows, N, d, (flags & DSS_AFF_NORMALIZE) ? 1 : 0);
  }
  DSS_CHECK_CUDA(cudaGetLastError());
  dim3 grid(cdiv(N, AT), cdiv(N, AT), B);
  LaunchScope scope(st, KC_AFFINITY);
  affinity_kernel<<<grid, 256, 0, st>>>(fn, img_max, color_counts, color_lambda, Wmat, N, d, ldw,
                                        (flags & DSS_AFF_THRESHOLD_AT_ZERO) ? 1 : 0);
  DSS_CHECK_CUDA(cudaGetLastError());
  return DSS_OK;
}

// out = epilogue(A[M,K] * Wt[N,K]^T + bias) with fp16 operands, fp32 accumulation in Tensor Memory.
//
// One CTA computes one 128x128 output tile:
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, 64-wide K slabs, STAGES-deep ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x128x16, kind::f16)
//   warps 2..5  : epilogue       (tcgen05.ld 32 lanes x 32 columns -> bias / GELU / residual -> global)
// Two such CTAs fit on one SM (96 KB smem, 128 TMEM columns each) so one tile's epilogue overlaps the other's
// main loop. Both operands are K-major, which is the native layout of activations [rows, features] and of
// torch Linear weights [out, in]; no transposes anywhere.
#include <math.h>

#include "common.cuh"

namespace dss {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16, STAGES = 3;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int B_TILE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 128;
// ring + 1024 B alignment slack + barriers/bias
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 1024;

struct EpiParams {
  void* out;
  const float* bias;
  const float* aux;
  int ldo;
  int rin, rout;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Row of the output buffer that GEMM row m maps to (or -1: skip).
template <int EPI>
__device__ __forceinline__ long long out_row(int m, const EpiParams& p) {
  if constexpr (EPI == DSS_EPI_PATCH_F32) {
    return (long long)(m / p.rin) * p.rout + (m % p.rin) + 1;
  } else if constexpr (EPI == DSS_EPI_DROPCLS_F32) {
    const int t = m % p.rin;
    return t == 0 ? -1 : (long long)(m / p.rin) * p.rout + t - 1;
  } else {
    return m;
  }
}

// Applies the epilogue to 32 consecutive columns [n, n+32) of one row and stores them.
template <int EPI>
__device__ __forceinline__ void epilogue_store(const float (&v)[32], int m, int n, const EpiParams& p) {
  const long long r = out_row<EPI>(m, p);
  if (r < 0) return;
  if constexpr (EPI == DSS_EPI_BIAS_F16 || EPI == DSS_EPI_BIAS_GELU_F16) {
    __half* o = reinterpret_cast<__half*>(p.out) + r * p.ldo + n;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float x[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) x[t] = (EPI == DSS_EPI_BIAS_GELU_F16) ? gelu_erf(v[j + t]) : v[j + t];
      uint4 q;
      q.x = pack_half2(x[0], x[1]);
      q.y = pack_half2(x[2], x[3]);
      q.z = pack_half2(x[4], x[5]);
      q.w = pack_half2(x[6], x[7]);
      *reinterpret_cast<uint4*>(o + j) = q;
    }
  } else {
    float* o = reinterpret_cast<float*>(p.out) + r * p.ldo + n;
    const float* aux = nullptr;
    if constexpr (EPI == DSS_EPI_PATCH_F32) aux = p.aux + (long long)((m % p.rin) + 1) * p.ldo + n;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 x = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      if constexpr (EPI == DSS_EPI_BIAS_RESID_F32) {
        const float4 y = *reinterpret_cast<const float4*>(o + j);
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
      }
      if constexpr (EPI == DSS_EPI_PATCH_F32) {
        const float4 y = __ldg(reinterpret_cast<const float4*>(aux + j));
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
      }
      *reinterpret_cast<float4*>(o + j) = x;
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M,
                        int N, int K, EpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment (the swizzle pattern is a function of address bits [7,10))
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  // barrier block layout: full[STAGES] | empty[STAGES] | tmem_full | tmem_ptr(u32) ... bias[128] at +128
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gbase + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));
  float* bias_s = reinterpret_cast<float*>(gbase + STAGES * STAGE_BYTES + 128);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;  // 0..127
    bias_s[t] = (n0 + t < N) ? __ldg(p.bias + n0 + t) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), STAGE_BYTES);
        const uint32_t sa = base + s * STAGE_BYTES;
        tma_load_2d(sa, &tmA, full_bar(s), kb * BK, m0);
        tma_load_2d(sa + A_TILE_BYTES, &tmB, full_bar(s), kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = base + s * STAGE_BYTES;
        const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advancing K inside the 128 B swizzle atom = advancing the start address by k*16 elements*2 B
          const uint64_t adesc = umma_desc_sw128(sa + k * UMMA_K * 2);
          const uint64_t bdesc = umma_desc_sw128(sb + k * UMMA_K * 2);
          umma_f16_ss(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // smem slot is free once these MMAs have consumed it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // epilogue warps 2..5: a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      if (m < M && n0 + c * 32 < N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias_s[c * 32 + j];
        epilogue_store<EPI>(v, m, n0 + c * 32, p);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// CUDA-core checker with the same epilogues (tests only).
template <int EPI>
__global__ void gemm_f16_simt_kernel(const __half* __restrict__ A, const __half* __restrict__ Wt, int M, int N, int K,
                                     EpiParams p) {
  // one thread = one row x 32 columns
  const int m = blockIdx.y * blockDim.y + threadIdx.y;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 32;
  if (m >= M || n >= N) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = __half2float(A[(size_t)m * K + k]);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(a, __half2float(Wt[(size_t)(n + j) * K + k]), v[j]);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] += p.bias[n + j];
  epilogue_store<EPI>(v, m, n, p);
}

// ---------------------------------------------------------------------------------------------------------------
// Host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !ptr) {
    set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// 2D fp16 row-major [rows, cols] tensor, box = 64 columns x 128 rows, 128 B swizzle, zero fill out of bounds.
int make_tmap_f16(CUtensorMap* tm, const void* ptr, int rows, int cols) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return DSS_ERR_CUDA;
  DSS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  DSS_REQUIRE(cols % 8 == 0, "TMA operand row pitch must be a multiple of 16 bytes (cols=%d)", cols);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d)", (int)r, rows, cols);
    return DSS_ERR_CUDA;
  }
  return DSS_OK;
}

template <int EPI>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const EpiParams& p,
                     cudaStream_t st, int kclass) {
  static bool attr_set = false;
  if (!attr_set) {
    DSS_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        GEMM_SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
  LaunchScope scope(st, kclass);
  gemm_f16_tcgen05_kernel<EPI><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(tmA, tmB, M, N, K, p);
  DSS_CHECK_CUDA(cudaGetLastError());
  return DSS_OK;
}

static int check_gemm_args(int M, int N, int K, int epi, const float* bias, const void* out, const float* aux,
                           int rin, int rout) {
  DSS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  DSS_REQUIRE(N % 32 == 0 && K % 8 == 0, "gemm: need N %% 32 == 0 and K %% 8 == 0 (N=%d K=%d)", N, K);
  DSS_REQUIRE(bias && out, "gemm: null bias/out");
  if (epi == DSS_EPI_PATCH_F32 || epi == DSS_EPI_DROPCLS_F32) {
    DSS_REQUIRE(rin > 0 && rout > 0 && M % rin == 0, "gemm: bad row remap rin=%d rout=%d M=%d", rin, rout, M);
    if (epi == DSS_EPI_PATCH_F32) DSS_REQUIRE(aux != nullptr, "gemm: patch epilogue needs aux (pos embed)");
  }
  return DSS_OK;
}

// Launch with pre-built tensor maps (used by the ViT forward, which caches them).
int gemm_f16_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, void* out, int M, int N, int K,
                int epi, const float* aux, int rin, int rout, cudaStream_t st, int kclass) {
  int rc = check_gemm_args(M, N, K, epi, bias, out, aux, rin, rout);
  if (rc) return rc;
  EpiParams p{out, bias, aux, N, rin, rout};
  switch (epi) {
    case DSS_EPI_BIAS_F16: return launch_tc<DSS_EPI_BIAS_F16>(tmA, tmB, M, N, K, p, st, kclass);
    case DSS_EPI_BIAS_GELU_F16: return launch_tc<DSS_EPI_BIAS_GELU_F16>(tmA, tmB, M, N, K, p, st, kclass);
    case DSS_EPI_BIAS_RESID_F32: return launch_tc<DSS_EPI_BIAS_RESID_F32>(tmA, tmB, M, N, K, p, st, kclass);
    case DSS_EPI_BIAS_F32: return launch_tc<DSS_EPI_BIAS_F32>(tmA, tmB, M, N, K, p, st, kclass);
    case DSS_EPI_PATCH_F32: return launch_tc<DSS_EPI_PATCH_F32>(tmA, tmB, M, N, K, p, st, kclass);
    case DSS_EPI_DROPCLS_F32: return launch_tc<DSS_EPI_DROPCLS_F32>(tmA, tmB, M, N, K, p, st, kclass);
  }
  set_error("gemm: unknown epilogue %d", epi);
  return DSS_ERR_BAD_ARG;
}

template <int EPI>
static int launch_simt(const void* A, const void* Wt, int M, int N, int K, const EpiParams& p, cudaStream_t st) {
  dim3 block(4, 32);
  dim3 grid(cdiv(N / 32, 4), cdiv(M, 32));
  LaunchScope scope(st, KC_MISC);
  gemm_f16_simt_kernel<EPI><<<grid, block, 0, st>>>(reinterpret_cast<const __half*>(A),
                                                    reinterpret_cast<const __half*>(Wt), M, N, K, p);
  DSS_CHECK_CUDA(cudaGetLastError());
  return DSS_OK;
}

}  // namespace dss

using namespace dss;

extern "C" int dss_op_gemm_f16(const void* A, const void* Wt, const float* bias, void* out, int M, int N, int K,
                               int epilogue, const float* aux, int rin, int rout, dss_stream_t stream) {
  int rc = check_gemm_args(M, N, K, epilogue, bias, out, aux, rin, rout);
  if (rc) return rc;
  DSS_REQUIRE(A && Wt, "gemm: null operand");
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_f16(&tmA, A, M, K))) return rc;
  if ((rc = make_tmap_f16(&tmB, Wt, N, K))) return rc;
  return gemm_f16_tc(tmA, tmB, bias, out, M, N, K, epilogue, aux, rin, rout, static_cast<cudaStream_t>(stream),
                     KC_GEMM_OTHER);
}

extern "C" int dss_op_gemm_f16_simt(const void* A, const void* Wt, const float* bias, void* out, int M, int N, int K,
                                    int epilogue, const float* aux, int rin, int rout, dss_stream_t stream) {
  int rc = check_gemm_args(M, N, K, epilogue, bias, out, aux, rin, rout);
  if (rc) return rc;
  DSS_REQUIRE(A && Wt, "gemm: null operand");
  EpiParams p{out, bias, aux, N, rin, rout};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case DSS_EPI_BIAS_F16: return launch_simt<DSS_EPI_BIAS_F16>(A, Wt, M, N, K, p, st);
    case DSS_EPI_BIAS_GELU_F16: return launch_simt<DSS_EPI_BIAS_GELU_F16>(A, Wt, M, N, K, p, st);
    case DSS_EPI_BIAS_RESID_F32: return launch_simt<DSS_EPI_BIAS_RESID_F32>(A, Wt, M, N, K, p, st);
    case DSS_EPI_BIAS_F32: return launch_simt<DSS_EPI_BIAS_F32>(A, Wt, M, N, K, p, st);
    case DSS_EPI_PATCH_F32: return launch_simt<DSS_EPI_PATCH_F32>(A, Wt, M, N, K, p, st);
    case DSS_EPI_DROPCLS_F32: return launch_simt<DSS_EPI_DROPCLS_F32>(A, Wt, M, N, K, p, st);
  }
  set_error("gemm: unknown epilogue %d", epilogue);
  return DSS_ERR_BAD_ARG;
}

// Multi-head attention (head dim 64) on the 5th-generation tensor cores: O = softmax(Q K^T / 8) V.
//
// One CTA = one (image, head, 128-query tile), one CTA per SM, software-pipelined flash-style loop over 128-key tiles:
//   warp 0      TMA producer: Q once, then K_j / V_j tiles straight out of the packed qkv activations
//               [B*T, 3d] (box 64 x 128, 128 B swizzle) through a 3-stage ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//                 S_j = Q K_j^T   (UMMA 128x128x16 x4, both operands K-major)            -> TMEM S[j & 1]
//                 O_j = P_j V_j   (UMMA 128x64x16  x8, A = P from smem, B = V MN-major)   -> TMEM O[j & 1]
//               S_{j+1} is issued BEFORE waiting for P_j, so the tensor core computes the next score tile while the
//               softmax warps work on the current one; S, O and P are all double buffered.
//   warps 2..9  softmax / accumulate: TWO threads per query row (tcgen05.ld 32x32b): warps 2..5 own keys 0..63 and
//               head dims 0..31, warps 6..9 keys 64..127 and head dims 32..63; the partial row maxima are exchanged
//               through shared memory. Running max / sum in base 2, P_j written as fp16 into the K-major 128 B-
//               swizzled smem tiles the PV MMA reads. The output is accumulated in REGISTERS one tile late
//               (O = O * 2^(m_{j-2} - m_{j-1}) + O_{j-1} while tile j is in flight), so the softmax warps never
//               wait for a tensor-core result that was issued in the same iteration.
// The kernel is bound by the MUFU pipe (128 x 128 exp2 per tile at 16 per clock per SM), not by the tensor pipe.
#include <math.h>

#include "common.cuh"

namespace dss {

int make_tmap_f16(CUtensorMap* tm, const void* ptr, int rows, int cols, int box_rows);

// one MUFU.EX2 (2 ulp), flushes denormal results to zero; exp2f would add range checks and fix-ups per element
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64, FA_THREADS = 320;   // TMA warp + MMA warp + 8 softmax warps
constexpr int FA_TILE = FA_BM * FA_D * 2;             // 16 KB: one [128 x 64] fp16 tile
constexpr int FA_KV_STAGES = 3;
constexpr int FA_SMEM = FA_TILE * (1 + 2 * FA_KV_STAGES + 4);  // Q | K ring | V ring | P[2] (2 atoms each) = 176 KB
constexpr int FA_TMEM_COLS = 512;
constexpr int FA_S_COL = 0, FA_O_COL = 256;   // S[2] at columns 0 / 128, O[2] at columns 256 / 320

__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int T, int heads) {
  extern __shared__ __align__(1024) uint8_t fa_smem[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * FA_KV_STAGES + 6];  // q_full | kv_full[] | kv_empty[] | s_full[2] | p_full[2] | o_full[2]
  __shared__ uint32_t tmem_ptr_s;
  __shared__ float xch[2][2][FA_BM];   // [tile parity][key half][row]: partial row maxima (and the final row sums)

  const uint32_t base = smem_u32(fa_smem);
  if ((base & 1023u) != 0) __trap();  // the 128 B swizzle pattern is a function of address bits [7,10)
  // P_j = 128 queries x 128 keys fp16 = two 64-key atoms of 16 KB; double buffered
  const uint32_t sQ = base, sK = base + FA_TILE, sV = sK + FA_KV_STAGES * FA_TILE, sP = sV + FA_KV_STAGES * FA_TILE;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0;
  auto kv_full = [&](int s) { return bar0 + 8u + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 8u + 8u * (FA_KV_STAGES + s); };
  auto s_full = [&](int i) { return bar0 + 8u + 8u * (2 * FA_KV_STAGES + i); };
  auto p_full = [&](int i) { return bar0 + 8u + 8u * (2 * FA_KV_STAGES + 2 + i); };
  auto o_full = [&](int i) { return bar0 + 8u + 8u * (2 * FA_KV_STAGES + 4 + i); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int d = heads * FA_D;
  const int q0 = qt * FA_BM;
  const int nt = (T + FA_BN - 1) / FA_BN;
  const int row0 = b * T;  // first row of this image in the [B*T, 3d] activation matrix

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 256);
      mbar_init(o_full(i), 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_ptr_s), FA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_s);

  // Producer and MMA warps run their loops with ALL lanes on warp-uniform values and let one elected lane issue: the
  // smem addresses, coordinates and descriptors then live in uniform registers and each UMMA / TMA issue is a couple
  // of instructions instead of a register-to-uniform "waterfall" loop.
  if (warp == 0) {
    const uint32_t uQ = __shfl_sync(0xffffffffu, sQ, 0), uK = __shfl_sync(0xffffffffu, sK, 0),
                   uV = __shfl_sync(0xffffffffu, sV, 0);
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, FA_TILE);
      tma_load_2d(uQ, &tmQKV, q_full, h * FA_D, row0 + q0);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % FA_KV_STAGES;
      const uint32_t ph = (j / FA_KV_STAGES) & 1;
      mbar_wait(kv_empty(s), ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full(s), 2 * FA_TILE);
        tma_load_2d(uK + s * FA_TILE, &tmQKV, kv_full(s), d + h * FA_D, row0 + j * FA_BN);
        tma_load_2d(uV + s * FA_TILE, &tmQKV, kv_full(s), 2 * d + h * FA_D, row0 + j * FA_BN);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = umma_idesc_f16(FA_BM, FA_BN);                 // A, B K-major
    constexpr uint32_t idesc_pv = umma_idesc_f16(FA_BM, FA_D) | (1u << 16);     // B (= V) MN-major
    const uint32_t uQ = __shfl_sync(0xffffffffu, sQ, 0), uK = __shfl_sync(0xffffffffu, sK, 0),
                   uV = __shfl_sync(0xffffffffu, sV, 0), uP = __shfl_sync(0xffffffffu, sP, 0),
                   utmem = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t qdesc = umma_desc_sw128(uQ);
    mbar_wait(q_full, 0);
    auto issue_s = [&](int j) {   // S[j & 1] = Q K_j^T
      const int s = j % FA_KV_STAGES;
      mbar_wait(kv_full(s), (j / FA_KV_STAGES) & 1);
      tc_fence_after();
      const uint64_t kdesc = umma_desc_sw128(uK + s * FA_TILE);
      const uint32_t acc = utmem + FA_S_COL + (j & 1) * FA_BN;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k)   // +32 B per 16-wide K step = +2 in the descriptor's address field
          umma_f16_ss(acc, qdesc + 2u * k, kdesc + 2u * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(s_full(j & 1));
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < nt; ++j) {
      // next score tile first: its S buffer was released by p_full(j-1), observed in the previous iteration
      if (j + 1 < nt) issue_s(j + 1);
      // O[j & 1] = P_j V_j once the softmax warps have published P_j (they consumed O_{j-2} before that)
      mbar_wait(p_full(j & 1), (j >> 1) & 1);
      tc_fence_after();
      const int s = j % FA_KV_STAGES;
      const uint64_t p0 = umma_desc_sw128(uP + (j & 1) * 2 * FA_TILE);              // keys 0..63 (K-major atom)
      const uint64_t p1 = umma_desc_sw128(uP + (j & 1) * 2 * FA_TILE + FA_TILE);    // keys 64..127
      const uint64_t vdesc = umma_desc_sw128(uV + s * FA_TILE);                     // +16 key rows = +2048 B = +128
      const uint32_t acc = utmem + FA_O_COL + (j & 1) * FA_D;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)
          umma_f16_ss(acc, (k < 4 ? p0 : p1) + 2u * (k & 3), vdesc + 128u * k, idesc_pv, k != 0 ? 1u : 0u);
        umma_commit(o_full(j & 1));
        umma_commit(kv_empty(s));
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int hs = (warp - 2) >> 2;       // 0: keys 0..63 / head dims 0..31, 1: keys 64..127 / head dims 32..63
    const int r = q * 32 + lane;          // query row inside the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float sc = 1.4426950408889634f * 0.125f;  // log2(e) / sqrt(64)
    float m_run = -INFINITY, l_run = 0.f;  // l_run: this thread's half of the row sum
    float corr_prev = 0.f;                 // rescale factor that goes with the not-yet-accumulated O_{j-1}
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    auto accumulate_o = [&](int jj, float corr) {   // o = o * corr + O_jj  (head dims [32 hs, 32 hs + 32))
      mbar_wait(o_full(jj & 1), (jj >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(lane_addr + FA_O_COL + (jj & 1) * FA_D + hs * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], corr, __uint_as_float(v[i]));
    };
    for (int j = 0; j < nt; ++j) {
      mbar_wait(s_full(j & 1), (j >> 1) & 1);
      tc_fence_after();
      const int nvalid = T - j * FA_BN - hs * 64;   // keys of this thread's half tile that exist
      const bool full_tile = nvalid >= 64;
      // this row inside the P atom of this half (P[j & 1] was last read by PV_{j-2}: complete, see accumulate_o below)
      const uint32_t prow = sP + ((j & 1) * 2 + hs) * FA_TILE + r * 128;
      const uint32_t scol = lane_addr + FA_S_COL + (j & 1) * FA_BN + hs * 64;
      // this thread's 64 scores are read from TMEM ONCE and stay in registers (TMEM -> RF bandwidth is the scarce
      // resource of this kernel after the MUFU pipe)
      uint32_t v[64];
      {
        uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        tmem_ld_32x32(scol, v0);
        tmem_ld_32x32(scol + 32, v1);
        tmem_ld_wait();
      }
      // row maximum over this half, exchanged with the thread owning the other half
      if (!full_tile) {   // keys beyond T (only in the last tile): score -inf -> probability exactly 0
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= nvalid) v[i] = 0xff800000u;
      }
      float mxp = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) mxp = fmaxf(mxp, __uint_as_float(v[i]));
      xch[j & 1][hs][r] = mxp;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float mx = fmaxf(m_run, fmaxf(mxp, xch[j & 1][hs ^ 1][r]));
      const float corr = ex2_approx((m_run - mx) * sc);   // first tile: exp2(-inf) = 0
      const float msc = mx * sc;
      m_run = mx;
      // P = 2^(s*c - m*c), partial row sum, fp16 P into the swizzled K-major tile of this half
      float rs = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {   // 8 chunks of 8 keys = 16 bytes each; chunk g of the row sits at slot g ^ (r % 8)
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const int i = g * 8 + e;
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -msc));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -msc));
          rs += p0 + p1;
          pk[e >> 1] = pack_half2(p0, p1);
        }
        const uint32_t addr = prow + (((g ^ (r & 7))) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                     : "memory");
      }
      l_run = l_run * corr + rs;
      fence_proxy_async_smem();   // generic-proxy writes of P -> visible to the tensor core (async proxy)
      tc_fence_before();          // order the TMEM reads of S_j (and of O_{j-2}) before the MMAs that overwrite them
      mbar_arrive(p_full(j & 1));
      // accumulate the PREVIOUS tile's P V product (issued one iteration ago, so normally already complete)
      if (j > 0) accumulate_o(j - 1, corr_prev);
      corr_prev = corr;
    }
    accumulate_o(nt - 1, corr_prev);
    // epilogue: total row sum = sum of the two halves; normalise, stage the [128 x 64] fp16 tile in the (now idle)
    // K/V ring, store coalesced
    xch[nt & 1][hs][r] = l_run;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float inv = 1.0f / (l_run + xch[nt & 1][hs ^ 1][r]);
    uint8_t* stage = fa_smem + FA_TILE;       // rows of 128 B + 16 B pad -> conflict-free row-per-thread writes
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 w;
      w.x = pack_half2(o[i + 0] * inv, o[i + 1] * inv);
      w.y = pack_half2(o[i + 2] * inv, o[i + 3] * inv);
      w.z = pack_half2(o[i + 4] * inv, o[i + 5] * inv);
      w.w = pack_half2(o[i + 6] * inv, o[i + 7] * inv);
      *reinterpret_cast<uint4*>(stage + r * 144 + hs * 64 + i * 2) = w;
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const int ew = warp - 2;
    __half* og = out + (long long)row0 * d + h * FA_D;
#pragma unroll 4
    for (int it = 0; it < 4; ++it) {
      const int rr = ew * 16 + it * 4 + (lane >> 3);    // 4 rows per warp instruction, 8 lanes x 16 B per row
      const int t = q0 + rr;
      if (t < T) {
        const uint4 w = *reinterpret_cast<const uint4*>(stage + rr * 144 + (lane & 7) * 16);
        *reinterpret_cast<uint4*>(og + (long long)t * d + (lane & 7) * 8) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

int launch_attention_tc(const void* qkv, void* out, int B, int T, int heads, cudaStream_t st) {
  DSS_REQUIRE(B > 0 && T > 0 && heads > 0, "attention: empty problem");
  static bool attr_set = false;
  if (!attr_set) {
    DSS_CHECK_CUDA(cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    attr_set = true;
  }
  CUtensorMap tm;
  int rc = make_tmap_f16(&tm, qkv, B * T, 3 * heads * FA_D, FA_BM);
  if (rc) return rc;
  dim3 grid(cdiv(T, FA_BM), heads, B);
  LaunchScope scope(st, KC_ATTENTION);
  attention_tcgen05_kernel<<<grid, FA_THREADS, FA_SMEM, st>>>(tm, reinterpret_cast<__half*>(out), T, heads);
  DSS_CHECK_CUDA(cudaGetLastError());
  return DSS_OK;
}

}  // namespace dss

extern "C" int dss_op_attention_tc_f16(const void* qkv, void* out, int B, int T, int heads, dss_stream_t stream) {
  return dss::launch_attention_tc(qkv, out, B, T, heads, static_cast<cudaStream_t>(stream));
}

// Multi-head attention (head dim 64) on the 5th-generation tensor cores: O = softmax(Q K^T / 8) V.
//
// One CTA = one (image, head, 128-query tile). Flash-style loop over 128-key tiles:
//   warp 0      TMA producer: Q once, then K_j / V_j tiles straight out of the packed qkv activations
//               [B*T, 3d] (box 64 x 128, 128 B swizzle) through a 2-stage ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//                 S   = Q K_j^T   (UMMA 128x128x16 x4, both operands K-major)            -> TMEM cols [0,128)
//                 O_j = P_j V_j   (UMMA 128x64x16  x8, A = P from smem, B = V MN-major)   -> TMEM cols [128,192)
//   warps 2..5  softmax / accumulate: thread = query row (tcgen05.ld 32x32b), running max / sum in base 2,
//               P_j written as fp16 into the K-major 128 B-swizzled smem tile the PV MMA reads, and the output
//               accumulated in REGISTERS: O = O * 2^(m_old - m_new) + O_j, so TMEM is never rescaled in place.
// Two CTAs are resident per SM (96 KB smem, 256 TMEM columns each): while one CTA's softmax warps are busy
// (MUFU-bound: 128x128 exp2 per tile) the other CTA's MMAs run.
#include <math.h>

#include "common.cuh"

namespace dss {

int make_tmap_f16(CUtensorMap* tm, const void* ptr, int rows, int cols, int box_rows);

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64, FA_THREADS = 192;
constexpr int FA_TILE = FA_BM * FA_D * 2;             // 16 KB: one [128 x 64] fp16 tile
constexpr int FA_SMEM = FA_TILE * (1 + 2 + 2 + 1);  // Q | K0 K1 | V0 V1 | P(keys 64..127) = 96 KB; P(keys 0..63) reuses K_j
constexpr int FA_TMEM_COLS = 256;
constexpr int FA_S_COL = 0, FA_O_COL = 128;

__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int T, int heads) {
  extern __shared__ __align__(1024) uint8_t fa_smem[];
  __shared__ __align__(8) uint64_t bars[9];  // q_full | kv_full[2] | kv_empty[2] | s_full | p_full | o_full | (pad)
  __shared__ uint32_t tmem_ptr_s;

  const uint32_t base = smem_u32(fa_smem);
  if ((base & 1023u) != 0) __trap();  // the 128 B swizzle pattern is a function of address bits [7,10)
  // P_j (128 queries x 128 keys fp16 = two 64-key atoms): atom 0 overwrites K_j, which is dead once S_j = Q K_j^T has
  // completed (s_full) and is not refilled before PV_j has completed (kv_empty); atom 1 has its own buffer.
  const uint32_t sQ = base, sK = base + FA_TILE, sV = base + 3 * FA_TILE, sP1 = base + 5 * FA_TILE;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, s_full = bar0 + 40, p_full = bar0 + 48, o_full = bar0 + 56;
  auto kv_full = [&](int s) { return bar0 + 8u + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 24u + 8u * s; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int d = heads * FA_D;
  const int q0 = qt * FA_BM;
  const int nt = (T + FA_BN - 1) / FA_BN;
  const int row0 = b * T;  // first row of this image in the [B*T, 3d] activation matrix

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, FA_TILE);
      tma_load_2d(sQ, &tmQKV, q_full, h * FA_D, row0 + q0);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(kv_empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(kv_full(s), 2 * FA_TILE);
        tma_load_2d(sK + s * FA_TILE, &tmQKV, kv_full(s), d + h * FA_D, row0 + j * FA_BN);
        tma_load_2d(sV + s * FA_TILE, &tmQKV, kv_full(s), 2 * d + h * FA_D, row0 + j * FA_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(FA_BM, FA_BN);                 // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_f16(FA_BM, FA_D) | (1u << 16);     // B (= V) MN-major
      mbar_wait(q_full, 0);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(kv_full(s), ph);
        tc_fence_after();
        // S = Q K^T. The S columns are free: softmax of tile j-1 finished reading them before p_full(j-1).
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k)
          umma_f16_ss(tmem_base + FA_S_COL, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(sK + s * FA_TILE + k * 32),
                      idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(s_full);
        // O_j = P_j V_j once the softmax warps have published P_j (and consumed O_{j-1})
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          const uint32_t patom = (k >> 2) ? sP1 : sK + s * FA_TILE;                        // 64-key atoms, K-major
          const uint64_t adesc = umma_desc_sw128(patom + (k & 3) * 32);
          const uint64_t bdesc = umma_desc_sw128(sV + s * FA_TILE + k * 16 * 128);          // 16 key rows per step
          umma_f16_ss(tmem_base + FA_O_COL, adesc, bdesc, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(s));
      }
    }
  } else {
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;          // query row inside the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float sc = 1.4426950408889634f * 0.125f;  // log2(e) / sqrt(64)
    float m_run = -INFINITY, l_run = 0.f;
    float o[FA_D];
#pragma unroll
    for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int nvalid = T - j * FA_BN;   // keys of this tile that exist (>= 128 except for the last tile)
      const bool full_tile = nvalid >= FA_BN;
      const uint32_t prow0 = sK + (j & 1) * FA_TILE + r * 128;   // this row inside P atom 0 (keys 0..63)
      const uint32_t prow1 = sP1 + r * 128;                      // ... and atom 1 (keys 64..127)
      // pass 1: row maximum (two 32-column TMEM loads in flight per wait)
      float mx = m_run;
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(lane_addr + FA_S_COL + c * 32, v0);
        tmem_ld_32x32(lane_addr + FA_S_COL + c * 32 + 32, v1);
        tmem_ld_wait();
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v0[i]));
            if (c * 32 + 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v1[i]));
          }
        }
      }
      const float corr = exp2f((m_run - mx) * sc);   // first tile: exp2(-inf) = 0
      const float msc = mx * sc;
      m_run = mx;
      // pass 2: P = 2^(s*c - m*c), row sum, fp16 P into the swizzled K-major tiles
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + FA_S_COL + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = exp2f(fmaf(__uint_as_float(v[i]), sc, -msc));
            const float p1 = exp2f(fmaf(__uint_as_float(v[i + 1]), sc, -msc));
            rs += p0 + p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < nvalid) ? exp2f(fmaf(__uint_as_float(v[i]), sc, -msc)) : 0.f;
            const float p1 = (c * 32 + i + 1 < nvalid) ? exp2f(fmaf(__uint_as_float(v[i + 1]), sc, -msc)) : 0.f;
            rs += p0 + p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
        }
        // keys [32c, 32c+32) = 4 chunks of 8 keys; chunk g of the row: atom g/8, slot (g%8) ^ (r%8)
        const uint32_t prow = (c < 2) ? prow0 : prow1;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int g = (c & 1) * 4 + t;
          const uint32_t addr = prow + (((g ^ (r & 7))) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[t * 4 + 0]), "r"(pk[t * 4 + 1]),
                       "r"(pk[t * 4 + 2]), "r"(pk[t * 4 + 3])
                       : "memory");
        }
      }
      l_run = l_run * corr + rs;
      fence_proxy_async_smem();   // generic-proxy writes of P -> visible to the tensor core (async proxy)
      tc_fence_before();          // order the TMEM reads of S before the next S MMA
      mbar_arrive(p_full);
      // O = O * corr + P_j V_j
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(lane_addr + FA_O_COL, v0);
        tmem_ld_32x32(lane_addr + FA_O_COL + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[i] = fmaf(o[i], corr, __uint_as_float(v0[i]));
          o[32 + i] = fmaf(o[32 + i], corr, __uint_as_float(v1[i]));
        }
      }
      tc_fence_before();          // O_j consumed before PV_{j+1} may overwrite it (ordered via p_full(j+1))
    }
    // epilogue: normalise, stage the [128 x 64] fp16 tile in the (now idle) P buffer, store coalesced
    const float inv = 1.0f / l_run;
    uint8_t* stage = fa_smem + FA_TILE;       // K/V ring is idle now; rows of 128 B + 16 B pad -> conflict-free writes
#pragma unroll
    for (int i = 0; i < FA_D; i += 8) {
      uint4 w;
      w.x = pack_half2(o[i + 0] * inv, o[i + 1] * inv);
      w.y = pack_half2(o[i + 2] * inv, o[i + 3] * inv);
      w.z = pack_half2(o[i + 4] * inv, o[i + 5] * inv);
      w.w = pack_half2(o[i + 6] * inv, o[i + 7] * inv);
      *reinterpret_cast<uint4*>(stage + r * 144 + i * 2) = w;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int ew = warp - 2;
    __half* og = out + (long long)row0 * d + h * FA_D;
#pragma unroll 4
    for (int it = 0; it < 8; ++it) {
      const int rr = ew * 32 + it * 4 + (lane >> 3);    // 4 rows per warp instruction, 8 lanes x 16 B per row
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

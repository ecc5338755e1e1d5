// Multi-head attention (head dim 64) on the 5th-generation tensor cores: O = softmax(Q K^T / 8) V.
//
// Persistent kernel, one CTA per SM. A work item is one (image, head, PAIR of 128-query tiles); the CTA walks its items
// with every ring / barrier phase carried across item boundaries, so the loads and the first score tile of the next
// item overlap the tail of the current one.
//   warp 0      TMA producer: Q tiles (double buffered per query tile) and K_j / V_j tiles straight out of the packed
//               qkv activations [B*T, 3d] (box 64 x 128, 128 B swizzle) through a 3-stage ring shared by both
//               query tiles
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, per query tile g in {A, B} and key tile j:
//                 S_g = Q_g K_j^T  (UMMA 128 x kc x 16, x4, both operands K-major)               -> TMEM S_g
//                 O_g = P_g V_j    (UMMA 128 x 64 x 16, x kc/16, A = P from smem, B = V MN-major) -> TMEM O_g[j & 1]
//               kc = 128 except in the last key tile, where it is the number of existing keys rounded up to 16
//               (T = 901: 16 instead of 128). S_g(j+1) is issued as soon as the softmax warps have read S_g(j), BEFORE
//               P_g V_j.
//   warps 2..5  softmax of query tile A, warps 6..9 of query tile B: ONE thread per query row (tcgen05.ld 32x32b), no
//               cross-thread exchange. Pass 1 reads the 128 scores for the row maximum (the second half read stays in
//               registers), pass 2 computes P = 2^(s c - m c) (MUFU.EX2), the row sum, and writes P as fp16 into the
//               K-major 128 B-swizzled smem tiles the PV MMA reads. The output is accumulated in REGISTERS one tile
//               late (o = o * 2^(m_{j-1} - m_j) + O_j after P_{j+1} has been published), so a softmax warp never waits
//               for a tensor-core result issued in the same iteration.
// The two softmax groups run out of phase (B's first tile is held back until A has finished its first pass 1), so one
// group's MUFU-bound exp phase overlaps the other's TMEM loads / maxima / fences / O accumulation. The kernel is bound
// by the MUFU pipe (128 x 128 exp2 per tile at 16 per clock per SM), not by the tensor pipe.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace dss {

int make_tmap_f16(CUtensorMap* tm, const void* ptr, int rows, int cols, int box_rows);

// one MUFU.EX2 (2 ulp), flushes denormal results to zero; exp2f would add range checks and fix-ups per element
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64, FA_THREADS = 320;   // TMA warp + MMA warp + 2 x 4 softmax warps
constexpr int FA_TILE = FA_BM * FA_D * 2;             // 16 KB: one [128 x 64] fp16 tile
constexpr int FA_KV_STAGES = 3;
// Q[g][2] | K ring | V ring | P[g] (two 64-key atoms each) = 14 tiles = 224 KB
constexpr int FA_SMEM = FA_TILE * (4 + 2 * FA_KV_STAGES + 4);
constexpr int FA_TMEM_COLS = 512;
constexpr int FA_S_COL = 0, FA_O_COL = 256;   // S_A, S_B at columns 0 / 128; O_g[i] at 256 + 128 g + 64 i
constexpr int FA_NBARS = 8 + 2 * FA_KV_STAGES + 10;

// 10 warps = 3 on some SM sub-partitions: 16 K registers / (3 x 32 threads) caps the kernel at 168 registers
__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int T, int heads,
                         int nq2, int total_items) {
  extern __shared__ __align__(1024) uint8_t fa_smem[];
  __shared__ __align__(8) uint64_t bars[FA_NBARS];
  __shared__ uint32_t tmem_ptr_s;

  const uint32_t base = smem_u32(fa_smem);
  if ((base & 1023u) != 0) __trap();  // the 128 B swizzle pattern is a function of address bits [7,10)
  const uint32_t sQ = base, sK = base + 4 * FA_TILE, sV = sK + FA_KV_STAGES * FA_TILE, sP = sV + FA_KV_STAGES * FA_TILE;
  const uint32_t bar0 = smem_u32(bars);
  auto q_full = [&](int g, int i) { return bar0 + 8u * (g * 2 + i); };
  auto q_empty = [&](int g, int i) { return bar0 + 8u * (4 + g * 2 + i); };
  auto kv_full = [&](int s) { return bar0 + 8u * (8 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (8 + FA_KV_STAGES + s); };
  auto s_full = [&](int g) { return bar0 + 8u * (8 + 2 * FA_KV_STAGES + g); };
  auto p_full = [&](int g) { return bar0 + 8u * (10 + 2 * FA_KV_STAGES + g); };
  auto o_full = [&](int g, int i) { return bar0 + 8u * (12 + 2 * FA_KV_STAGES + g * 2 + i); };
  auto s_free = [&](int g) { return bar0 + 8u * (16 + 2 * FA_KV_STAGES + g); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * FA_D;
  const int nt = (T + FA_BN - 1) / FA_BN;
  const int kc_last = (T - (nt - 1) * FA_BN + 15) & ~15;   // key columns of the last tile that are worth computing

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 4; ++i) {
      mbar_init(q_full(i >> 1, i & 1), 1);
      mbar_init(q_empty(i >> 1, i & 1), 1);
      mbar_init(o_full(i >> 1, i & 1), 1);
    }
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 128);
      mbar_init(s_free(g), 128);
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
    int s = 0;
    uint32_t ph = 0;   // K/V ring position, carried across items
    int qi = 0;        // item counter of this CTA
    for (int w = blockIdx.x; w < total_items; w += gridDim.x, ++qi) {
      const int qp = w % nq2, bh = w / nq2;
      const int h = bh % heads, row0 = (bh / heads) * T;   // first row of this image in the [B*T, 3d] matrix
      const int qb = qi & 1;
      const uint32_t qph = (qi >> 1) & 1;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        mbar_wait(q_empty(g, qb), qph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full(g, qb), FA_TILE);
          tma_load_2d(uQ + (g * 2 + qb) * FA_TILE, &tmQKV, q_full(g, qb), h * FA_D, row0 + (2 * qp + g) * FA_BM);
        }
        __syncwarp();
      }
      for (int j = 0; j < nt; ++j) {
        mbar_wait(kv_empty(s), ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(kv_full(s), 2 * FA_TILE);
          tma_load_2d(uK + s * FA_TILE, &tmQKV, kv_full(s), d + h * FA_D, row0 + j * FA_BN);
          tma_load_2d(uV + s * FA_TILE, &tmQKV, kv_full(s), 2 * d + h * FA_D, row0 + j * FA_BN);
        }
        __syncwarp();
        if (++s == FA_KV_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk0 = umma_idesc_f16(FA_BM, 0);                    // A, B K-major; N filled in per tile
    constexpr uint32_t idesc_pv = umma_idesc_f16(FA_BM, FA_D) | (1u << 16);     // B (= V) MN-major
    const uint32_t uQ = __shfl_sync(0xffffffffu, sQ, 0), uK = __shfl_sync(0xffffffffu, sK, 0),
                   uV = __shfl_sync(0xffffffffu, sV, 0), uP = __shfl_sync(0xffffffffu, sP, 0),
                   utmem = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int my_items = (total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int m_total = my_items * nt;   // key-tile steps of this CTA, over all of its items
    // Event loop: each query tile g has a cursor for its next score tile (needs S_g released by the softmax warps)
    // and one for its next P V product (needs P_g published); whichever is ready is issued.
    int ns[2] = {0, 0}, sj[2] = {0, 0}, sq[2] = {0, 0}, ss[2] = {0, 0};     // score cursor: step, key tile, item, ring stage
    uint32_t sph[2] = {0, 0};
    int np[2] = {0, 0}, pj[2] = {0, 0}, ps[2] = {0, 0};                      // P V cursor
    uint32_t idle = 0;
    while (np[0] < m_total || np[1] < m_total) {
      bool progress = false;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        // ---- S_g(n) = Q_g K_j^T: S_g is free once the softmax warps have read S_g(n-1) for the last time
        // (all probes are non-blocking: a blocking wait for K_j here could starve the other query tile's P V product,
        // which is what releases the ring slot K_j is waiting for)
        if (ns[g] < m_total && (ns[g] == 0 || mbar_test_all(s_free(g), (ns[g] - 1) & 1)) &&
            mbar_test_all(kv_full(ss[g]), sph[g]) &&
            (sj[g] != 0 || mbar_test_all(q_full(g, sq[g] & 1), (sq[g] >> 1) & 1))) {
          tc_fence_after();
          const uint64_t qdesc = umma_desc_sw128(uQ + (g * 2 + (sq[g] & 1)) * FA_TILE);
          const uint64_t kdesc = umma_desc_sw128(uK + ss[g] * FA_TILE);
          const int kc = sj[g] == nt - 1 ? kc_last : FA_BN;
          const uint32_t idesc = idesc_qk0 | (static_cast<uint32_t>(kc >> 3) << 17);
          const uint32_t acc = utmem + FA_S_COL + g * FA_BN;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < FA_D / 16; ++k)   // +32 B per 16-wide K step = +2 in the descriptor's address field
              umma_f16_ss(acc, qdesc + 2u * k, kdesc + 2u * k, idesc, k != 0 ? 1u : 0u);
            umma_commit(s_full(g));
            if (sj[g] == nt - 1) umma_commit(q_empty(g, sq[g] & 1));   // last score tile of the item: Q_g may be replaced
          }
          __syncwarp();
          ++ns[g];
          if (++sj[g] == nt) { sj[g] = 0; ++sq[g]; }
          if (++ss[g] == FA_KV_STAGES) { ss[g] = 0; sph[g] ^= 1u; }
          progress = true;
        }
        // ---- O_g[n & 1] = P_g V_j once the softmax warps have published P_g(n) (they consumed O_g(n-2) before that)
        if (np[g] < m_total && mbar_test_all(p_full(g), np[g] & 1)) {
          tc_fence_after();
          const int n = np[g];
          const int ksteps = (pj[g] == nt - 1 ? kc_last : FA_BN) >> 4;
          const uint64_t vdesc = umma_desc_sw128(uV + ps[g] * FA_TILE);   // +16 key rows = +2048 B = +128
          const uint64_t p0 = umma_desc_sw128(uP + g * 2 * FA_TILE);              // keys 0..63 (K-major atom)
          const uint64_t p1 = umma_desc_sw128(uP + g * 2 * FA_TILE + FA_TILE);    // keys 64..127
          const uint32_t acc = utmem + FA_O_COL + g * 2 * FA_D + (n & 1) * FA_D;
          const bool release = np[g ^ 1] > n;   // the other query tile has already issued its P V_j: K_j / V_j are done
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < FA_BN / 16; ++k)
              if (k < ksteps)
                umma_f16_ss(acc, (k < 4 ? p0 : p1) + 2u * (k & 3), vdesc + 128u * k, idesc_pv, k != 0 ? 1u : 0u);
            umma_commit(o_full(g, n & 1));
            if (release) umma_commit(kv_empty(ps[g]));
          }
          __syncwarp();
          ++np[g];
          if (++pj[g] == nt) pj[g] = 0;
          if (++ps[g] == FA_KV_STAGES) ps[g] = 0;
          progress = true;
        }
      }
      if (progress) idle = 0;
      else if (++idle > (1u << 26)) __trap();   // protocol bug -> trap instead of a hang
    }
  } else {
    const int g = (warp - 2) >> 2;        // query tile of the pair
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;          // query row inside the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_col = lane_addr + FA_S_COL + g * FA_BN;
    const uint32_t o_col = lane_addr + FA_O_COL + g * 2 * FA_D;
    const uint32_t prow = sP + g * 2 * FA_TILE + r * 128;   // this row inside the first P atom; second atom + FA_TILE
    const float sc = 1.4426950408889634f * 0.125f;  // log2(e) / sqrt(64)
    uint32_t m = 0;   // key-tile step counter of this CTA (all barrier phases derive from it)
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      const int qp = w % nq2, bh = w / nq2;
      const int h = bh % heads, row0 = (bh / heads) * T;
      const int q0 = (2 * qp + g) * FA_BM;
      const bool dead = q0 >= T;   // odd number of query tiles: nothing to do for B in the last pair
      float m_run = -INFINITY, l_run = 0.f;
      float corr_prev = 0.f;       // rescale factor that goes with the not-yet-accumulated O_{j-1}
      float o[FA_D];
#pragma unroll
      for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
      auto accumulate_o = [&](uint32_t mm, float corr) {   // o = o * corr + O(mm)
        uint32_t t[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(o_col + (mm & 1) * FA_D + c * 32, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], corr, __uint_as_float(t[i]));
        }
      };
      for (int j = 0; j < nt; ++j, ++m) {
        if (m == 0 && g == 1) asm volatile("bar.sync 3, 256;" ::: "memory");   // start half a period behind tile A
        mbar_wait(s_full(g), m & 1);
        tc_fence_after();
        if (!dead) {
          // one key tile; FULL = all 128 key columns exist (straight-line code, no masks), otherwise the last tile
          auto tile = [&](auto full_c) {
            constexpr bool FULL = decltype(full_c)::value;
            const int nvalid = FULL ? FA_BN : T - j * FA_BN;      // existing keys of this tile
            const int kc = FULL ? FA_BN : kc_last;                // key columns the MMAs cover
            uint32_t v[64];
            uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
            uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
            auto load_half = [&](int hf) {   // scores of keys [64 hf, 64 hf + 64); keys beyond T -> -inf -> P exactly 0
              tmem_ld_32x32(s_col + hf * 64, v0);
              tmem_ld_32x32(s_col + hf * 64 + 32, v1);
              tmem_ld_wait();
              if (!FULL) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                  if (hf * 64 + i >= nvalid) v[i] = 0xff800000u;
              }
            };
            // ---- pass 1: row maximum. Keys 64..127 first, then keys 0..63, which stay in registers for pass 2
            float mx = m_run;
            if (FULL || kc > 64) {
              load_half(1);
#pragma unroll
              for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            load_half(0);
#pragma unroll
            for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            if (!FULL && kc <= 64) {   // nothing more to read from S_g: let the next score tile start
              tc_fence_before();
              mbar_arrive(s_free(g));
            }
            if (m == 0 && g == 0) asm volatile("bar.arrive 3, 256;" ::: "memory");
            const float corr = ex2_approx((m_run - mx) * sc);   // first tile: exp2(-inf) = 0
            const float msc = mx * sc;
            m_run = mx;
            // P_g is single buffered: P_g V_{j-1} must have finished reading it (issued long ago: no stall in practice)
            if (j > 0) mbar_wait(o_full(g, (m - 1) & 1), ((m - 1) >> 1) & 1);
            // ---- pass 2: P = 2^(s*c - m*c), row sum, fp16 P into the swizzled K-major atoms
            float rs = 0.f;
            auto exp_store = [&](uint32_t pr, int cols) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {   // 8 chunks of 8 keys = 16 bytes each; chunk c of the row sits at slot c ^ (r % 8)
                if (FULL || c * 8 < cols) {
                  uint32_t pk[4];
#pragma unroll
                  for (int e = 0; e < 8; e += 2) {
                    const int i = c * 8 + e;
                    const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -msc));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -msc));
                    rs += p0 + p1;
                    pk[e >> 1] = pack_half2(p0, p1);
                  }
                  const uint32_t addr = pr + (((c ^ (r & 7))) << 4);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]),
                               "r"(pk[2]), "r"(pk[3])
                               : "memory");
                }
              }
            };
            exp_store(prow, kc);   // keys 0..63 (or the first kc of them)
            if (FULL || kc > 64) {
              load_half(1);
              tc_fence_before();
              mbar_arrive(s_free(g));   // last read of S_g: S_g(j+1) is computed while the second half is exponentiated
              exp_store(prow + FA_TILE, kc - 64);
            }
            l_run = l_run * corr + rs;
            fence_proxy_async_smem();   // generic-proxy writes of P -> visible to the tensor core (async proxy)
            tc_fence_before();          // order the TMEM reads of O_g(j-2) before the MMA that overwrites it
            mbar_arrive(p_full(g));
            // accumulate the PREVIOUS tile's P V product (complete, see the o_full wait above)
            if (j > 0) {
              tc_fence_after();
              accumulate_o(m - 1, corr_prev);
            }
            corr_prev = corr;
          };
          if (j < nt - 1 || (T & (FA_BN - 1)) == 0) tile(std::true_type{});
          else tile(std::false_type{});
        } else {
          tc_fence_before();
          mbar_arrive(s_free(g));
          mbar_arrive(p_full(g));
        }
      }
      if (!dead) {
        const uint32_t ml = m - 1;   // last tile of this item
        mbar_wait(o_full(g, ml & 1), (ml >> 1) & 1);
        tc_fence_after();
        accumulate_o(ml, corr_prev);
        tc_fence_before();
        const float inv = 1.0f / l_run;
        const int t = q0 + r;
        if (t < T) {   // each thread owns one full 128-byte output row
          uint4* og = reinterpret_cast<uint4*>(out + (long long)(row0 + t) * d + h * FA_D);
#pragma unroll
          for (int i = 0; i < FA_D; i += 8) {
            uint4 wv;
            wv.x = pack_half2(o[i + 0] * inv, o[i + 1] * inv);
            wv.y = pack_half2(o[i + 2] * inv, o[i + 3] * inv);
            wv.z = pack_half2(o[i + 4] * inv, o[i + 5] * inv);
            wv.w = pack_half2(o[i + 6] * inv, o[i + 7] * inv);
            og[i >> 3] = wv;
          }
        }
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
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    DSS_CHECK_CUDA(cudaGetDevice(&dev));
    DSS_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int nq2 = cdiv(cdiv(T, FA_BM), 2);   // pairs of 128-query tiles per (image, head)
  const int total = B * heads * nq2;
  const int grid = total < sm_count ? total : sm_count;   // persistent: one CTA per SM
  LaunchScope scope(st, KC_ATTENTION);
  attention_tcgen05_kernel<<<grid, FA_THREADS, FA_SMEM, st>>>(tm, reinterpret_cast<__half*>(out), T, heads, nq2, total);
  DSS_CHECK_CUDA(cudaGetLastError());
  return DSS_OK;
}

}  // namespace dss

extern "C" int dss_op_attention_tc_f16(const void* qkv, void* out, int B, int T, int heads, dss_stream_t stream) {
  return dss::launch_attention_tc(qkv, out, B, T, heads, static_cast<cudaStream_t>(stream));
}

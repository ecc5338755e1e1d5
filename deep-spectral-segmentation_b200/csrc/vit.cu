// DINO ViT feature extractor: weight packing, positional-embedding interpolation, forward orchestration.
// Replaces utils.get_model + model.get_intermediate_layers + the qkv forward hook of the reference
// (extract/extract_utils.py:40-50, extract/extract.py:49-53,82-98).
//
// Per block:  LN1 -> [tcgen05 GEMM] qkv (f16) -> tcgen05 flash attention -> [tcgen05 GEMM + residual] proj
//             LN2 -> [tcgen05 GEMM + GELU] fc1 (f16) -> [tcgen05 GEMM + residual] fc2
// The residual stream stays fp32; GEMM operands are fp16 (11-bit significand, same as TF32), accumulation fp32.
// The block the reference hooks is pruned to LN1 + the K third of its qkv projection, CLS row dropped in the
// GEMM epilogue, which writes the fp32 features directly in the reference's [B, N, d] layout.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace dss {

int make_tmap_f16(CUtensorMap* tm, const void* ptr, int rows, int cols, int box_rows);
int gemm_tile_n(int N);
int make_tmap_out(CUtensorMap* tm, const void* ptr, int rows, int cols, int is_f32);
int gemm_f16_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmC, const float* bias, void* out,
                int M, int N, int K, int epi, const float* aux, int rin, int rout, cudaStream_t st, int kclass, int bn);
int launch_im2col(const uint8_t* img, void* patches, int B, int H, int W, int P, cudaStream_t st);
int launch_cls_row(float* x, const float* cls, const float* pos, int B, int T, int d, cudaStream_t st);
int launch_layernorm(const float* x, const float* g, const float* b, void* y, int M, int d, float eps, cudaStream_t st);
int launch_attention_tc(const void* qkv, void* out, int B, int T, int heads, cudaStream_t st);
int gemm_ln_tile_n(int N);
int gemm_ln_f16_tc(const CUtensorMap& tmB, const CUtensorMap& tmC, const float* x, const float* gamma, const float* beta,
                   const float* bias, int M, int N, float eps, bool gelu, cudaStream_t st, int kclass);

// y[b, :] = LayerNorm(x[b * row_stride, :]) * gamma + beta in fp32 (one warp per row): the final norm of the CLS token
__global__ void __launch_bounds__(128)
layernorm_rows_f32_kernel(const float* __restrict__ x, long long row_stride, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ y, int rows, int d, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * row_stride;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += xr[i];
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
  for (int i = lane; i < d; i += 32) { const float a = xr[i] - mean; q = fmaf(a, a, q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  for (int i = lane; i < d; i += 32) y[(long long)row * d + i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
}

__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}

struct BlockW {
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *qkv_b, *proj_b, *fc1_b, *fc2_b;
  __half *qkv_w, *proj_w, *fc1_w, *fc2_w;
  CUtensorMap tm_qkv, tm_k, tm_proj, tm_fc1, tm_fc2;
  CUtensorMap tm_qkv_ln, tm_fc1_ln;   // weight maps of the LayerNorm-fused kernel (its own tile width)
};

}  // namespace dss

struct dss_vit {
  dss_vit_config cfg;
  bool loaded = false;
  void* arena = nullptr;  // one device allocation holding every packed weight
  size_t arena_bytes = 0;
  __half* patch_w = nullptr;
  float *patch_b = nullptr, *cls = nullptr;
  float *norm_w = nullptr, *norm_b = nullptr;   // final LayerNorm (only needed by dss_vit_forward_cls)
  CUtensorMap tm_patch;
  std::vector<dss::BlockW> blocks;
  std::vector<float> pos_host;                        // [1 + grid0^2, d] fp32 host copy
  std::map<std::pair<int, int>, float*> pos_cache;    // (Hp, Wp) -> device [T, d]
  std::vector<std::pair<int, int>> pos_order;         // insertion order: the cache is bounded (oldest grid evicted)
};

namespace dss {

// ---- bicubic interpolation of the positional grid exactly as torch F.interpolate(mode='bicubic',
// align_corners=False, scale_factor=s) computes it: src = (dst + 0.5) / s - 0.5, A = -0.75, border clamp.
static inline double cubic1(double x, double A) { return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0; }
static inline double cubic2(double x, double A) { return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A; }

static void interp_axis(int in, int out, double scale_factor, std::vector<int>& idx, std::vector<float>& w) {
  idx.resize(out);
  w.resize((size_t)out * 4);
  const double A = -0.75;
  const float rscale = (float)(1.0 / scale_factor);  // torch keeps the scale in fp32 for fp32 inputs
  for (int o = 0; o < out; ++o) {
    const float real = rscale * ((float)o + 0.5f) - 0.5f;
    const float fl = floorf(real);
    const double t = (double)(real - fl);
    idx[o] = (int)fl;
    w[o * 4 + 0] = (float)cubic2(t + 1.0, A);
    w[o * 4 + 1] = (float)cubic1(t, A);
    w[o * 4 + 2] = (float)cubic1(1.0 - t, A);
    w[o * 4 + 3] = (float)cubic2(2.0 - t, A);
  }
  (void)in;
}

// pos_src [1 + g*g, d] -> pos_out [1 + Hp*Wp, d]; host memory, pure CPU (runs once per distinct image shape)
static void interp_pos_host(const float* src, int g, int d, int Hp, int Wp, float* pos) {
  std::copy(src, src + d, pos);  // CLS slot untouched
  if (Hp == g && Wp == g) {
    std::copy(src + d, src + (size_t)(g * g + 1) * d, pos + d);
    return;
  }
  // upstream: scale_factor = ((Hp + 0.1) / sqrt(N0), (Wp + 0.1) / sqrt(N0)); output size floor(g * scale) == Hp, Wp
  std::vector<int> iy, ix;
  std::vector<float> wy, wx;
  interp_axis(g, Hp, ((double)Hp + 0.1) / (double)g, iy, wy);
  interp_axis(g, Wp, ((double)Wp + 0.1) / (double)g, ix, wx);
  auto clampi = [g](int v) { return v < 0 ? 0 : (v >= g ? g - 1 : v); };
  for (int y = 0; y < Hp; ++y)
    for (int x = 0; x < Wp; ++x) {
      float* o = pos + ((size_t)(y * Wp + x) + 1) * d;
      for (int c = 0; c < d; ++c) o[c] = 0.f;
      for (int a = 0; a < 4; ++a) {
        const int sy = clampi(iy[y] - 1 + a);
        for (int b = 0; b < 4; ++b) {
          const int sx = clampi(ix[x] - 1 + b);
          const float wgt = wy[y * 4 + a] * wx[x * 4 + b];
          const float* s = src + ((size_t)(sy * g + sx) + 1) * d;
          for (int c = 0; c < d; ++c) o[c] += wgt * s[c];
        }
      }
    }
}

static int get_pos(dss_vit* h, int Hp, int Wp, cudaStream_t st, float** out) {
  auto key = std::make_pair(Hp, Wp);
  auto it = h->pos_cache.find(key);
  if (it != h->pos_cache.end()) {
    *out = it->second;
    return DSS_OK;
  }
  const int d = h->cfg.dim, g = h->cfg.grid0, T = Hp * Wp + 1;
  std::vector<float> pos((size_t)T * d);
  interp_pos_host(h->pos_host.data(), g, d, Hp, Wp, pos.data());
  float* dev = nullptr;
  DSS_CHECK_CUDA(cudaMalloc(&dev, (size_t)T * d * sizeof(float)));
  DSS_CHECK_CUDA(cudaMemcpyAsync(dev, pos.data(), (size_t)T * d * sizeof(float), cudaMemcpyHostToDevice, st));
  DSS_CHECK_CUDA(cudaStreamSynchronize(st));  // `pos` is a pageable temporary
  // bounded: data sets with hundreds of distinct image sizes must not grow the cache without limit. Eviction happens
  // after the stream synchronisation above, so no kernel that was given the evicted pointer is still running.
  constexpr size_t kMaxPosGrids = 64;
  if (h->pos_order.size() >= kMaxPosGrids) {
    auto old = h->pos_order.front();
    h->pos_order.erase(h->pos_order.begin());
    auto oit = h->pos_cache.find(old);
    if (oit != h->pos_cache.end()) {
      cudaFree(oit->second);
      h->pos_cache.erase(oit);
    }
  }
  h->pos_cache[key] = dev;
  h->pos_order.push_back(key);
  *out = dev;
  return DSS_OK;
}

struct VitWs {
  __half* patches;  // [B*Np, 3P^2]
  float* x;         // [B*T, d]     residual stream
  __half* xn;       // [B*T, d]     LayerNorm output (GEMM A operand)
  __half* attn;     // [B*T, d]     attention output (GEMM A operand)
  __half* qkv;      // [B*T, 3d]
  __half* hid;      // [B*T, 4d]
  size_t total;
};

static VitWs carve(const dss_vit_config& c, int B, int H, int W, void* base) {
  const int Np = (H / c.patch) * (W / c.patch), T = Np + 1, d = c.dim;
  const size_t M = (size_t)B * T;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  VitWs w;
  w.patches = reinterpret_cast<__half*>(take((size_t)B * Np * 3 * c.patch * c.patch * 2));
  w.x = reinterpret_cast<float*>(take(M * d * 4));
  w.xn = reinterpret_cast<__half*>(take(M * d * 2));
  w.attn = reinterpret_cast<__half*>(take(M * d * 2));
  w.qkv = reinterpret_cast<__half*>(take(M * 3 * d * 2));
  w.hid = reinterpret_cast<__half*>(take(M * (size_t)c.mlp_ratio * d * 2));
  w.total = off;
  return w;
}

// mode 0: K projection of block n_full (features), 1: residual stream [B, T, d], 2: final-norm CLS token [B, d]
static int vit_run(dss_vit* h, const uint8_t* img, int B, int H, int W, int n_full, int mode, float* out, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const bool k_proj = mode == 0;
  DSS_REQUIRE(h && h->loaded, "vit: weights not loaded");
  DSS_REQUIRE(img && out && ws, "vit: null pointer");
  const dss_vit_config& c = h->cfg;
  const int P = c.patch, d = c.dim;
  const int Hp = H / P, Wp = W / P;
  DSS_REQUIRE(B > 0 && Hp > 0 && Wp > 0, "vit: image %dx%d smaller than one %d-pixel patch", H, W, P);
  DSS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "vit: workspace must be 256-byte aligned");
  const int Np = Hp * Wp, T = Np + 1, M = B * T;
  VitWs w = carve(c, B, H, W, ws);
  if (ws_bytes < w.total) {
    set_error("vit: workspace too small (%zu < %zu)", ws_bytes, w.total);
    return DSS_ERR_WORKSPACE;
  }
  int rc;
  float* pos = nullptr;
  if ((rc = get_pos(h, Hp, Wp, st, &pos))) return rc;

  CUtensorMap tm_patches, tm_xn, tm_attn, tm_hid;
  const int Kp = 3 * P * P;
  if ((rc = make_tmap_f16(&tm_patches, w.patches, B * Np, Kp, 128))) return rc;
  if ((rc = make_tmap_f16(&tm_xn, w.xn, M, d, 128))) return rc;
  if ((rc = make_tmap_f16(&tm_attn, w.attn, M, d, 128))) return rc;
  if ((rc = make_tmap_f16(&tm_hid, w.hid, M, c.mlp_ratio * d, 128))) return rc;
  // output maps of the TMA-store epilogues: residual stream (fp32, reduce-add), qkv and MLP hidden (fp16)
  CUtensorMap tm_x_out, tm_qkv_out, tm_hid_out;
  if ((rc = make_tmap_out(&tm_x_out, w.x, M, d, 1))) return rc;
  if ((rc = make_tmap_out(&tm_qkv_out, w.qkv, M, 3 * d, 0))) return rc;
  if ((rc = make_tmap_out(&tm_hid_out, w.hid, M, c.mlp_ratio * d, 0))) return rc;

  // tokens: x[b, 1+n, :] = patch_embed + pos ; x[b, 0, :] = cls + pos[0]
  if ((rc = launch_im2col(img, w.patches, B, H, W, P, st))) return rc;
  if ((rc = gemm_f16_tc(tm_patches, h->tm_patch, nullptr, h->patch_b, w.x, B * Np, d, Kp, DSS_EPI_PATCH_F32, pos, Np, T, st,
                        KC_GEMM_PATCH, 128)))
    return rc;
  if ((rc = launch_cls_row(w.x, h->cls, pos, B, T, d, st))) return rc;

  // K = 384 (ViT-S): LayerNorm is fused into the A-operand producer of the qkv / fc1 GEMMs (gemm_ln.cu); the
  // stand-alone LayerNorm kernel remains for ViT-B (a 128 x 768 fp16 panel does not fit next to the weight ring)
  static const bool fused_ln_env = [] { const char* e = getenv("DSS_VIT_FUSED_LN"); return !e || atoi(e) != 0; }();
  const bool fused_ln = fused_ln_env && d == 384;
  for (int l = 0; l < n_full; ++l) {
    const BlockW& bw = h->blocks[l];
    if (fused_ln) {
      if ((rc = gemm_ln_f16_tc(bw.tm_qkv_ln, tm_qkv_out, w.x, bw.ln1_w, bw.ln1_b, bw.qkv_b, M, 3 * d, c.ln_eps, false, st,
                               KC_GEMM_QKV)))
        return rc;
    } else {
      if ((rc = launch_layernorm(w.x, bw.ln1_w, bw.ln1_b, w.xn, M, d, c.ln_eps, st))) return rc;
      if ((rc = gemm_f16_tc(tm_xn, bw.tm_qkv, &tm_qkv_out, bw.qkv_b, w.qkv, M, 3 * d, d, DSS_EPI_BIAS_F16, nullptr, 0, 0, st,
                            KC_GEMM_QKV, gemm_tile_n(3 * d))))
        return rc;
    }
    if ((rc = launch_attention_tc(w.qkv, w.attn, B, T, c.heads, st))) return rc;
    if ((rc = gemm_f16_tc(tm_attn, bw.tm_proj, &tm_x_out, bw.proj_b, w.x, M, d, d, DSS_EPI_BIAS_RESID_F32, nullptr, 0, 0, st,
                          KC_GEMM_PROJ, gemm_tile_n(d))))
      return rc;
    if (fused_ln) {
      if ((rc = gemm_ln_f16_tc(bw.tm_fc1_ln, tm_hid_out, w.x, bw.ln2_w, bw.ln2_b, bw.fc1_b, M, c.mlp_ratio * d, c.ln_eps, true,
                               st, KC_GEMM_FC1)))
        return rc;
    } else {
      if ((rc = launch_layernorm(w.x, bw.ln2_w, bw.ln2_b, w.xn, M, d, c.ln_eps, st))) return rc;
      if ((rc = gemm_f16_tc(tm_xn, bw.tm_fc1, &tm_hid_out, bw.fc1_b, w.hid, M, c.mlp_ratio * d, d, DSS_EPI_BIAS_GELU_F16,
                            nullptr, 0, 0, st, KC_GEMM_FC1, gemm_tile_n(c.mlp_ratio * d))))
        return rc;
    }
    if ((rc = gemm_f16_tc(tm_hid, bw.tm_fc2, &tm_x_out, bw.fc2_b, w.x, M, d, c.mlp_ratio * d, DSS_EPI_BIAS_RESID_F32, nullptr, 0,
                          0, st, KC_GEMM_FC2, gemm_tile_n(d))))
      return rc;
  }
  if (k_proj) {
    const BlockW& bw = h->blocks[n_full];
    if ((rc = launch_layernorm(w.x, bw.ln1_w, bw.ln1_b, w.xn, M, d, c.ln_eps, st))) return rc;
    // K third of the qkv projection (weight rows [d, 2d)), CLS rows dropped: out[b, n, :] == qkv[b, 1+n, d:2d]
    if ((rc = gemm_f16_tc(tm_xn, bw.tm_k, nullptr, bw.qkv_b + d, out, M, d, d, DSS_EPI_DROPCLS_F32, nullptr, T, Np, st,
                          KC_GEMM_KPROJ, 128)))
      return rc;
  } else if (mode == 2) {
    // upstream VisionTransformer.forward: x = norm(x); return x[:, 0]  (LayerNorm is row-wise: only the CLS rows)
    DSS_REQUIRE(h->norm_w && h->norm_b, "vit: the final LayerNorm weights (norm.weight / norm.bias) were not loaded");
    LaunchScope scope(st, KC_LAYERNORM);
    layernorm_rows_f32_kernel<<<cdiv(B, 4), 128, 0, st>>>(w.x, (long long)T * d, h->norm_w, h->norm_b, out, B, d, c.ln_eps);
    DSS_CHECK_CUDA(cudaGetLastError());
  } else {
    DSS_CHECK_CUDA(cudaMemcpyAsync(out, w.x, (size_t)M * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return DSS_OK;
}

}  // namespace dss

using namespace dss;

extern "C" int dss_vit_create(const dss_vit_config* cfg, dss_vit_t** out) {
  DSS_REQUIRE(cfg && out, "vit_create: null pointer");
  DSS_REQUIRE(cfg->dim == 384 || cfg->dim == 768, "vit_create: dim must be 384 or 768 (got %d)", cfg->dim);
  DSS_REQUIRE(cfg->heads > 0 && cfg->dim == cfg->heads * 64, "vit_create: head dim must be 64 (dim=%d heads=%d)",
              cfg->dim, cfg->heads);
  DSS_REQUIRE(cfg->patch == 8 || cfg->patch == 16, "vit_create: patch must be 8 or 16 (got %d)", cfg->patch);
  DSS_REQUIRE(cfg->depth > 0 && cfg->mlp_ratio > 0 && cfg->grid0 > 0, "vit_create: bad depth/mlp_ratio/grid0");
  dss_vit* h = new dss_vit();
  h->cfg = *cfg;
  *out = h;
  return DSS_OK;
}

extern "C" void dss_vit_destroy(dss_vit_t* h) {
  if (!h) return;
  for (auto& kv : h->pos_cache) cudaFree(kv.second);
  if (h->arena) cudaFree(h->arena);
  delete h;
}

extern "C" int dss_vit_load_weights(dss_vit_t* h, const dss_vit_weights* w, dss_stream_t stream) {
  DSS_REQUIRE(h && w && w->blocks, "vit_load_weights: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dss_vit_config& c = h->cfg;
  const size_t d = c.dim, P = c.patch, hid = (size_t)c.mlp_ratio * d, Kp = 3 * P * P;
  // arena layout: f16 weights first (1 KB aligned for TMA), then fp32 vectors
  size_t off = 0;
  auto reserve = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 1024);
    return o;
  };
  const size_t o_patch_w = reserve(d * Kp * 2), o_patch_b = reserve(d * 4), o_cls = reserve(d * 4);
  const size_t o_norm_w = reserve(d * 4), o_norm_b = reserve(d * 4);
  struct Off { size_t qkv_w, proj_w, fc1_w, fc2_w, ln1_w, ln1_b, ln2_w, ln2_b, qkv_b, proj_b, fc1_b, fc2_b; };
  std::vector<Off> bo(c.depth);
  for (int l = 0; l < c.depth; ++l) {
    bo[l].qkv_w = reserve(3 * d * d * 2); bo[l].proj_w = reserve(d * d * 2);
    bo[l].fc1_w = reserve(hid * d * 2); bo[l].fc2_w = reserve(d * hid * 2);
    bo[l].ln1_w = reserve(d * 4); bo[l].ln1_b = reserve(d * 4); bo[l].ln2_w = reserve(d * 4); bo[l].ln2_b = reserve(d * 4);
    bo[l].qkv_b = reserve(3 * d * 4); bo[l].proj_b = reserve(d * 4); bo[l].fc1_b = reserve(hid * 4); bo[l].fc2_b = reserve(d * 4);
  }
  if (h->arena && h->arena_bytes < off) { cudaFree(h->arena); h->arena = nullptr; }
  if (!h->arena) {
    DSS_CHECK_CUDA(cudaMalloc(&h->arena, off));
    h->arena_bytes = off;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(h->arena);
  auto cvt = [&](const float* src, size_t o, size_t n) -> int {
    DSS_REQUIRE(src, "vit_load_weights: null weight pointer");
    LaunchScope scope(st, KC_MISC);
    f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, reinterpret_cast<__half*>(base + o), n);
    DSS_CHECK_CUDA(cudaGetLastError());
    return DSS_OK;
  };
  auto cpy = [&](const float* src, size_t o, size_t n) -> int {
    DSS_REQUIRE(src, "vit_load_weights: null weight pointer");
    DSS_CHECK_CUDA(cudaMemcpyAsync(base + o, src, n * 4, cudaMemcpyDeviceToDevice, st));
    return DSS_OK;
  };
  int rc;
  if ((rc = cvt(w->patch_w, o_patch_w, d * Kp))) return rc;
  if ((rc = cpy(w->patch_b, o_patch_b, d))) return rc;
  if ((rc = cpy(w->cls_token, o_cls, d))) return rc;
  h->patch_w = reinterpret_cast<__half*>(base + o_patch_w);
  h->patch_b = reinterpret_cast<float*>(base + o_patch_b);
  h->cls = reinterpret_cast<float*>(base + o_cls);
  h->norm_w = h->norm_b = nullptr;
  if (w->norm_w && w->norm_b) {
    if ((rc = cpy(w->norm_w, o_norm_w, d))) return rc;
    if ((rc = cpy(w->norm_b, o_norm_b, d))) return rc;
    h->norm_w = reinterpret_cast<float*>(base + o_norm_w);
    h->norm_b = reinterpret_cast<float*>(base + o_norm_b);
  }
  if ((rc = make_tmap_f16(&h->tm_patch, h->patch_w, (int)d, (int)Kp, 128 / 2))) return rc;
  h->blocks.resize(c.depth);
  for (int l = 0; l < c.depth; ++l) {
    const dss_vit_block_weights& s = w->blocks[l];
    BlockW& b = h->blocks[l];
    if ((rc = cvt(s.qkv_w, bo[l].qkv_w, 3 * d * d))) return rc;
    if ((rc = cvt(s.proj_w, bo[l].proj_w, d * d))) return rc;
    if ((rc = cvt(s.fc1_w, bo[l].fc1_w, hid * d))) return rc;
    if ((rc = cvt(s.fc2_w, bo[l].fc2_w, d * hid))) return rc;
    if ((rc = cpy(s.ln1_w, bo[l].ln1_w, d))) return rc;
    if ((rc = cpy(s.ln1_b, bo[l].ln1_b, d))) return rc;
    if ((rc = cpy(s.ln2_w, bo[l].ln2_w, d))) return rc;
    if ((rc = cpy(s.ln2_b, bo[l].ln2_b, d))) return rc;
    if ((rc = cpy(s.qkv_b, bo[l].qkv_b, 3 * d))) return rc;
    if ((rc = cpy(s.proj_b, bo[l].proj_b, d))) return rc;
    if ((rc = cpy(s.fc1_b, bo[l].fc1_b, hid))) return rc;
    if ((rc = cpy(s.fc2_b, bo[l].fc2_b, d))) return rc;
    b.qkv_w = reinterpret_cast<__half*>(base + bo[l].qkv_w);
    b.proj_w = reinterpret_cast<__half*>(base + bo[l].proj_w);
    b.fc1_w = reinterpret_cast<__half*>(base + bo[l].fc1_w);
    b.fc2_w = reinterpret_cast<__half*>(base + bo[l].fc2_w);
    b.ln1_w = reinterpret_cast<float*>(base + bo[l].ln1_w); b.ln1_b = reinterpret_cast<float*>(base + bo[l].ln1_b);
    b.ln2_w = reinterpret_cast<float*>(base + bo[l].ln2_w); b.ln2_b = reinterpret_cast<float*>(base + bo[l].ln2_b);
    b.qkv_b = reinterpret_cast<float*>(base + bo[l].qkv_b); b.proj_b = reinterpret_cast<float*>(base + bo[l].proj_b);
    b.fc1_b = reinterpret_cast<float*>(base + bo[l].fc1_b); b.fc2_b = reinterpret_cast<float*>(base + bo[l].fc2_b);
    if ((rc = make_tmap_f16(&b.tm_qkv, b.qkv_w, (int)(3 * d), (int)d, gemm_tile_n((int)(3 * d)) / 2))) return rc;
    if ((rc = make_tmap_f16(&b.tm_k, b.qkv_w + d * d, (int)d, (int)d, 128 / 2))) return rc;
    if ((rc = make_tmap_f16(&b.tm_proj, b.proj_w, (int)d, (int)d, gemm_tile_n((int)d) / 2))) return rc;
    if ((rc = make_tmap_f16(&b.tm_fc1, b.fc1_w, (int)hid, (int)d, gemm_tile_n((int)hid) / 2))) return rc;
    if ((rc = make_tmap_f16(&b.tm_fc2, b.fc2_w, (int)d, (int)hid, gemm_tile_n((int)d) / 2))) return rc;
    if (d == 384) {
      if ((rc = make_tmap_f16(&b.tm_qkv_ln, b.qkv_w, (int)(3 * d), (int)d, gemm_ln_tile_n((int)(3 * d)) / 2))) return rc;
      if ((rc = make_tmap_f16(&b.tm_fc1_ln, b.fc1_w, (int)hid, (int)d, gemm_ln_tile_n((int)hid) / 2))) return rc;
    }
  }
  // host copy of the positional grid for interpolation; drop stale interpolations
  const size_t npos = ((size_t)c.grid0 * c.grid0 + 1) * d;
  h->pos_host.resize(npos);
  DSS_REQUIRE(w->pos_embed, "vit_load_weights: null pos_embed");
  DSS_CHECK_CUDA(cudaMemcpyAsync(h->pos_host.data(), w->pos_embed, npos * 4, cudaMemcpyDeviceToHost, st));
  DSS_CHECK_CUDA(cudaStreamSynchronize(st));
  for (auto& kv : h->pos_cache) cudaFree(kv.second);
  h->pos_cache.clear();
  h->pos_order.clear();
  h->loaded = true;
  return DSS_OK;
}

extern "C" size_t dss_vit_workspace_bytes(const dss_vit_t* h, int B, int H, int W) {
  if (!h || B <= 0 || H < h->cfg.patch || W < h->cfg.patch) return 0;
  return carve(h->cfg, B, H, W, nullptr).total;
}

extern "C" int dss_vit_forward_k(dss_vit_t* h, const uint8_t* images_u8, int B, int H, int W, int which_block,
                                 float* k_out, void* ws, size_t ws_bytes, dss_stream_t stream) {
  DSS_REQUIRE(h, "vit_forward_k: null handle");
  const int depth = h->cfg.depth;
  DSS_REQUIRE(which_block >= -depth && which_block < depth, "vit_forward_k: which_block %d out of range", which_block);
  const int blk = which_block < 0 ? which_block + depth : which_block;
  return vit_run(h, images_u8, B, H, W, blk, 0, k_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int dss_vit_forward_tokens(dss_vit_t* h, const uint8_t* images_u8, int B, int H, int W, int n_blocks,
                                      float* x_out, void* ws, size_t ws_bytes, dss_stream_t stream) {
  DSS_REQUIRE(h, "vit_forward_tokens: null handle");
  DSS_REQUIRE(n_blocks >= 0 && n_blocks <= h->cfg.depth, "vit_forward_tokens: n_blocks %d out of range", n_blocks);
  return vit_run(h, images_u8, B, H, W, n_blocks, 1, x_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int dss_vit_forward_cls(dss_vit_t* h, const uint8_t* images_u8, int B, int H, int W, float* cls_out, void* ws,
                                   size_t ws_bytes, dss_stream_t stream) {
  DSS_REQUIRE(h, "vit_forward_cls: null handle");
  return vit_run(h, images_u8, B, H, W, h->cfg.depth, 2, cls_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int dss_vit_pos_embed(dss_vit_t* h, int Hp, int Wp, float* out, dss_stream_t stream) {
  DSS_REQUIRE(h && h->loaded && out, "vit_pos_embed: handle not loaded / null out");
  DSS_REQUIRE(Hp > 0 && Wp > 0, "vit_pos_embed: empty grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* pos = nullptr;
  int rc = get_pos(h, Hp, Wp, st, &pos);
  if (rc) return rc;
  DSS_CHECK_CUDA(cudaMemcpyAsync(out, pos, ((size_t)Hp * Wp + 1) * h->cfg.dim * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
  return DSS_OK;
}

extern "C" int dss_pos_embed_interp_host(const float* pos_embed_host, int grid0, int d, int Hp, int Wp,
                                         float* out_host) {
  DSS_REQUIRE(pos_embed_host && out_host && grid0 > 0 && d > 0 && Hp > 0 && Wp > 0, "pos_embed_interp_host: bad args");
  interp_pos_host(pos_embed_host, grid0, d, Hp, Wp, out_host);
  return DSS_OK;
}

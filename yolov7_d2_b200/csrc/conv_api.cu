// C-ABI entry points for the convolution family (forward, data gradient, weight gradient, weight packing).
#include "conv_gemm.cuh"
#include "host_common.cuh"
#include "wgrad_gemm.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace yb {

static bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YB200_CONV_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool use_v1_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YB200_CONV_KERNEL");
    v = (e && strcmp(e, "v1") == 0) ? 1 : 0;
  }
  return v == 1;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------------
static bool use_staged_epilogue() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YB200_CONV_STAGED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int BN, int BK>
static int launch_conv_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, dim3 grid, int stages,
                            cudaStream_t st, const yb200_act* out_act) {
  using Cfg = ConvGemmCfg<BN, BK>;
  const bool ext = p.epi_mode >= EPI_BF16_AFFINE;  // ConvNeXt / transformer epilogues live in their own instantiations
  if constexpr (BN == 32 || BN == 64) {
    // narrow column tiles: staged epilogue (swizzled smem tile -> TMA store, BatchNorm statistics on the tensor core)
    const bool plain = (p.epi_mode == EPI_F16_STATS || p.epi_mode == EPI_F16 || p.epi_mode == EPI_BF16) && p.addend == nullptr && p.num_bnseg == 0;
    if (use_staged_epilogue() && !use_v1_kernel() && plain && out_act != nullptr && p.cout % BN == 0 && p.out_mh == 1 && p.out_mw == 1 && p.out_sc == 1) {
      const int m_tiles = grid.x, n_tiles = grid.y;
      const int tw = 1 << p.log_tw, th = 1 << p.log_th, tn = 128 >> (p.log_tw + p.log_th);
      CUtensorMap tmOut;
      int rc = make_act_map(&tmOut, *out_act, false, BN, tw, th, tn);
      if (rc) return rc;
      const int fixed = 2 * 128 * BN * 2 + 2048;  // two staged tiles + the two constant ones operands
      const int budget = 110 * 1024 - 1024 - fixed;
      int slots_kb = budget / Cfg::kStageBytes;
      const int num_kb = p.num_taps * p.cin_blocks;
      int kbs = 1;
      if (BK <= 32)
        for (int t = 1; t <= num_kb; ++t)
          if (num_kb % t == 0 && t * BK <= 144 && 2 * t <= slots_kb) kbs = t;
      int pst = slots_kb / kbs;
      if (pst > kMaxStagesP) pst = kMaxStagesP;
      if (pst < 2) pst = 2;
      const int smem = fixed + pst * kbs * Cfg::kStageBytes + 1024;
      YB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_staged_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int groups = (2 * sm_count()) / n_tiles;
      if (groups < 1) groups = 1;
      if (groups > m_tiles) groups = m_tiles;
      launch_k(conv_gemm_staged_kernel<BN, BK>, groups * n_tiles, kConvThreadsP, smem, st, tmA, tmB, tmOut, p, pst, kbs, n_tiles, m_tiles, out_act->c_off);
      YB_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  if (use_v1_kernel() && p.num_bnseg == 0) {  // one tile per CTA (kept for A/B comparison)
    static PerDevice<int> max_set_dev(0);
  int& max_set = max_set_dev.cur();
    const int smem = stages * Cfg::kStageBytes + 1024;
    if (smem > max_set) {
      YB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      max_set = smem;
    }
    launch_k(conv_gemm_kernel<BN, BK>, grid, kConvThreads, smem, st, tmA, tmB, p, stages);
    YB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if constexpr (BN == 256) {
    if (use_pair_kernel() && p.epi_mode != EPI_F32_BIAS && p.num_bnseg == 0 && p.num_phases == 0) {
      // CTA pairs: 2 x 128 pixels x 256 channels per UMMA, one CTA per SM, 32 KB per stage and CTA at BLOCK_K 64
      const int m_tiles = grid.x, n_tiles = grid.y;
      constexpr int stage_bytes = 2 * 128 * BK * 2;
      int pst = (212 * 1024) / stage_bytes;
      if (pst > kMaxStagesP) pst = kMaxStagesP;
      const int smem = pst * stage_bytes + 1024;
      static PerDevice<int> max_set_pair_dev(0);
  int& max_set_pair = max_set_pair_dev.cur();
      if (smem > max_set_pair) {
        YB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_pair_kernel<BK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        YB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_pair_kernel<BK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        max_set_pair = smem;
      }
      const int pair_tiles = (m_tiles + 1) / 2;
      int groups = (sm_count() / 2) / n_tiles;
      if (groups < 1) groups = 1;
      if (groups > pair_tiles) groups = pair_tiles;
      if (ext)
        launch_k(conv_gemm_pair_kernel<BK, true>, 2 * groups * n_tiles, kConvThreadsP, smem, st, tmA, tmB, p, pst, n_tiles, m_tiles);
      else
        launch_k(conv_gemm_pair_kernel<BK, false>, 2 * groups * n_tiles, kConvThreadsP, smem, st, tmA, tmB, p, pst, n_tiles, m_tiles);
      YB_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  // persistent kernel: two CTAs per SM (2 x 2 x BN TMEM columns <= 512; one CTA for BN = 256), ring as deep as the CTA's share
  // of shared memory allows
  const int phases = p.num_phases == 4 ? 4 : 1;
  const int m_tiles = grid.x * phases, n_tiles = grid.y;  // work items of one column tile (pixel tiles x output-parity phases)
  const int occ = BN == 256 ? 1 : 2;
  // fp32 rows with channel stride 1 and an odd pitch ([B, A, 85]): chunks go through a per-warp transpose scratch behind the ring
  const bool xpose = p.epi_mode == EPI_F32_BIAS && p.out_sc == 1 && BN <= 128 && p.num_bnseg == 0;
  const int budget = (occ == 1 ? 216 : 110) * 1024 - 1024 - (xpose ? kXposeBytes : 0);
  int slots_kb = budget / Cfg::kStageBytes;  // k-blocks that fit in the ring
  // narrow layers (BLOCK_K 16 / 32) would spend their time on mbarrier round trips: put several k-blocks (up to 144
  // channels-taps) behind one barrier, keeping at least two ring slots
  const int num_kb = phases == 4 ? p.cin_blocks : p.num_taps * p.cin_blocks;  // phases: 1 / 2 / 2 / 4 (or 1 each) taps -- slots must divide all
  int kbs = 1;
  if (BK <= 32)
    for (int t = 1; t <= num_kb; ++t)
      if (num_kb % t == 0 && t * BK <= 144 && 2 * t <= slots_kb) kbs = t;
  int pst = slots_kb / kbs;
  if (pst > kMaxStagesP) pst = kMaxStagesP;
  if (pst < 2) pst = 2;
  const int smem = pst * kbs * Cfg::kStageBytes + 1024 + (xpose ? kXposeBytes : 0);
  static PerDevice<int> max_set_p_dev(0);
  int& max_set_p = max_set_p_dev.cur();
  if (smem > max_set_p) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_persistent_kernel<BN, BK, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_persistent_kernel<BN, BK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    max_set_p = smem;
  }
  int groups = (occ * sm_count()) / n_tiles;
  if (groups < 1) groups = 1;
  if (groups > m_tiles) groups = m_tiles;
  if (phases == 4 && groups > 1 && groups % 2 == 0) --groups;  // odd stride through the work items: every CTA cycles through all four phases (1 / 2 / 2 / 4 taps)
  if (p.num_bnseg > 0) {  // data gradient with fused BatchNorm-backward statistics
    if constexpr (BN <= 128) {
      YB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_persistent_kernel<BN, BK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      launch_k(conv_gemm_persistent_kernel<BN, BK, 2>, groups * n_tiles, kConvThreadsP, smem, st, tmA, tmB, p, pst, kbs, n_tiles, m_tiles);
      YB_CHECK_CUDA(cudaGetLastError());
      return 0;
    } else {
      return fail(YB200_ERR_UNSUPPORTED, "fused BatchNorm-backward statistics need a column tile <= 128 (gradient tensors of < 256 channels)");
    }
  }
  ConvGemmParams pp = p;
  pp.xpose = xpose ? 1 : 0;
  if (ext)
    launch_k(conv_gemm_persistent_kernel<BN, BK, 1>, groups * n_tiles, kConvThreadsP, smem, st, tmA, tmB, pp, pst, kbs, n_tiles, m_tiles);
  else
    launch_k(conv_gemm_persistent_kernel<BN, BK, 0>, groups * n_tiles, kConvThreadsP, smem, st, tmA, tmB, pp, pst, kbs, n_tiles, m_tiles);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_conv(int bn, int bk, const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, dim3 grid,
                       int stages, cudaStream_t st, const yb200_act* out_act = nullptr) {
#define YB_CASE(BN, BK) \
  if (bn == BN && bk == BK) return launch_conv_inst<BN, BK>(tmA, tmB, p, grid, stages, st, out_act);
  YB_CASE(16, 16) YB_CASE(16, 32) YB_CASE(16, 64)
  YB_CASE(32, 16) YB_CASE(32, 32) YB_CASE(32, 64)
  YB_CASE(64, 16) YB_CASE(64, 32) YB_CASE(64, 64)
  YB_CASE(128, 16) YB_CASE(128, 32) YB_CASE(128, 64)
  YB_CASE(256, 16) YB_CASE(256, 32) YB_CASE(256, 64)
#undef YB_CASE
  return fail(YB200_ERR_UNSUPPORTED, "no conv_gemm instantiation for BLOCK_N=%d BLOCK_K=%d", bn, bk);
}

// pipeline depth: at most kMaxStages, never more than the K loop, and shallow enough that two CTAs fit in one SM's shared
// memory (one CTA computes one tile: a second resident CTA hides its prologue / epilogue behind the other's MMAs)
static int pick_stages(int bn, int bk, int num_kb) {
  const int stage_bytes = (128 + bn) * bk * 2;
  int st = num_kb < kMaxStages ? num_kb : kMaxStages;
  while (st > 2 && 2 * (st * stage_bytes + 2048) > 227 * 1024) --st;
  return st;
}
static int pick_block_k(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : (c % 16 == 0 ? 16 : 0)); }
// 256-wide column tiles halve the activation (A operand) traffic per MMA: the 3x3 layers are L2-bandwidth bound at 128
static int pick_block_n(int c) {
  static int allow256 = -1;
  if (allow256 < 0) {
    const char* e = getenv("YB200_CONV_BN256");
    allow256 = (e && e[0] == '0') ? 0 : 1;
  }
  if (c > 256 && c % 256 == 128) return 128;  // e.g. 384 = 3 x 128: no half-empty 256-wide tile
  if (c >= 256 && allow256 && !use_v1_kernel()) return 256;
  return c > 64 ? 128 : (c > 32 ? 64 : (c > 16 ? 32 : 16));
}

static int check_act(const yb200_act* a, const char* name) {
  YB_REQUIRE(a != nullptr && a->ptr != nullptr, YB200_ERR_INVALID, "%s: null view", name);
  YB_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0 && a->c > 0, YB200_ERR_INVALID, "%s: empty extent", name);
  YB_REQUIRE(a->c % 8 == 0 && a->c_pitch % 8 == 0 && a->c_off % 8 == 0 && a->c_off + a->c <= a->c_pitch, YB200_ERR_INVALID,
             "%s: channels (c=%d pitch=%d off=%d) must be multiples of 8 with off+c<=pitch", name, a->c, a->c_pitch, a->c_off);
  return 0;
}

// fill the forward-style tap table (reads input pixel  stride*o + k - pad)
static int fill_fwd_taps(ConvTap* taps, const yb200_act& x, int ksize, int stride, int k_per_tap) {
  int nt = 0;
  if (ksize == 1) {
    taps[nt++] = ConvTap{x.c_off, 0, 0, 0, 0};
  } else if (ksize == 2) {  // 2x2 stride 2, no padding: input pixel (2*o + kh, 2*o + kw) = (row parity kh, column parity kw) of cell o
    for (int kh = 0; kh < 2; ++kh)
      for (int kw = 0; kw < 2; ++kw) taps[nt++] = ConvTap{kw * x.c_pitch + x.c_off, 0, kh, 0, (kh * 2 + kw) * k_per_tap};
  } else {
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        ConvTap t{};
        if (stride == 1) {
          t.c0 = x.c_off; t.dw = kw - 1; t.p = 0; t.dh = kh - 1;
        } else {  // input row 2*o + kh - 1: kh=0 -> (parity 1, o-1), kh=1 -> (0, o), kh=2 -> (1, o)
          t.p = (kh == 1) ? 0 : 1; t.dh = (kh == 0) ? -1 : 0;
          const int pw = (kw == 1) ? 0 : 1;
          t.dw = (kw == 0) ? -1 : 0;
          t.c0 = pw * x.c_pitch + x.c_off;
        }
        t.kb = (kh * 3 + kw) * k_per_tap;
        taps[nt++] = t;
      }
  }
  return nt;
}

static void set_out_view(ConvGemmParams& p, const yb200_act& o) {
  p.out = static_cast<__nv_bfloat16*>(o.ptr) + o.c_off;
  p.out_sw = o.c_pitch;
  p.out_sh = 1LL * o.c_pitch * o.w;
  p.out_sn = 1LL * o.c_pitch * o.w * o.h;
  p.out_sc = 1;
  p.out_mh = 1; p.out_ph = 0; p.out_mw = 1; p.out_pw = 0;
}

static void set_tiles(ConvGemmParams& p, int n, int h, int w) {
  choose_tile(n, h, w, 128, &p.log_tw, &p.log_th);
  const int tw = 1 << p.log_tw, th = 1 << p.log_th, tn = 128 >> (p.log_tw + p.log_th);
  p.tiles_w = ceil_div(w, tw); p.tiles_h = ceil_div(h, th); p.tiles_n = ceil_div(n, tn);
  p.n_valid = n; p.h_valid = h; p.w_valid = w;
}

}  // namespace yb

using namespace yb;

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, const float* __restrict__ cout_scale, int cout, int cin, int taps, int cout_pad,
                                        int cin_pad, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  pdl_sync();
  const long long total = 1LL * cout_pad * taps * cin_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = static_cast<int>(i % cin_pad);
    const int t = static_cast<int>((i / cin_pad) % taps);
    const int co = static_cast<int>(i / (1LL * cin_pad * taps));
    float v = (co < cout && ci < cin) ? w[(1LL * co * cin + ci) * taps + t] : 0.f;
    if (cout_scale != nullptr && co < cout) v *= cout_scale[co];
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    if (wf) wf[i] = b;
    if (wd) wd[(1LL * ci * taps + t) * cout_pad + co] = b;
  }
}

static int pack_weight_impl(const float* w_oihw, const float* cout_scale, int cout, int cin, int ksize, int cout_pad, int cin_pad, void* w_fwd,
                            void* w_dgrad, void* stream) {
  YB_REQUIRE(w_oihw && (w_fwd || w_dgrad), YB200_ERR_INVALID, "pack_conv_weight: null pointer");
  YB_REQUIRE(cout > 0 && cin > 0 && (ksize >= 1 && ksize <= 3) && cout_pad >= cout && cin_pad >= cin, YB200_ERR_INVALID,
             "pack_conv_weight: bad sizes cout=%d cin=%d k=%d pads=%d,%d", cout, cin, ksize, cout_pad, cin_pad);
  const long long total = 1LL * cout_pad * ksize * ksize * cin_pad;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 4096));
  launch_k(pack_conv_weight_kernel, blocks, 256, 0, as_stream(stream), w_oihw, cout_scale, cout, cin, ksize * ksize, cout_pad, cin_pad,
                                                                 static_cast<__nv_bfloat16*>(w_fwd),
                                                                 static_cast<__nv_bfloat16*>(w_dgrad));
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// every convolution of a plan in ONE launch: the per-layer kernels are a few microseconds of work each, so ~70 launches per step are pure
// launch latency at the head of the step.  `table` (device memory, built once per plan) lists the layers; `prefix[i]` = padded elements of
// layers 0..i-1 (prefix[n] = total), so a thread finds its layer by binary search.
constexpr int kPackMaxLayers = 256;
__global__ void __launch_bounds__(256)
pack_conv_weights_batched_kernel(const yb200_pack_desc* __restrict__ table, const long long* __restrict__ prefix, int n) {
  pdl_sync();
  // Work unit = one ROW of a packed operand, one warp per row: forward rows (co, tap) run over ci, data-gradient rows (ci, tap) over co, so both
  // outputs are written with consecutive 2-byte stores; the fp32 source is gathered (9 M parameters: L2 resident).  Row r of the launch belongs
  // to layer l with rows_before[l] <= r: the row prefix is rebuilt per block in shared memory (n <= 256 layers), one binary search per row.
  __shared__ int s_rows[kPackMaxLayers + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < n; ++l) {
      s_rows[l] = acc;
      const yb200_pack_desc d = table[l];
      const int taps = d.ksize * d.ksize;
      acc += (d.w_fwd ? d.cout_pad * taps : 0) + (d.w_dgrad ? d.cin_pad * taps : 0);
    }
    s_rows[n] = acc;
  }
  __syncthreads();
  const int total_rows = s_rows[n];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < total_rows; r += warps) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_rows[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const yb200_pack_desc d = table[lo];
    const int taps = d.ksize * d.ksize;
    int q = r - s_rows[lo];
    const int fwd_rows = d.w_fwd ? d.cout_pad * taps : 0;
    if (q < fwd_rows) {  // forward operand [cout_pad][taps][cin_pad]
      const int co = q / taps, t = q - co * taps;
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(d.w_fwd) + static_cast<size_t>(q) * d.cin_pad;
      const float* src = d.w_oihw + static_cast<size_t>(co) * d.cin * taps + t;
      for (int ci = lane; ci < d.cin_pad; ci += 32)
        dst[ci] = __float2bfloat16_rn((co < d.cout && ci < d.cin) ? src[static_cast<size_t>(ci) * taps] : 0.f);
    } else {             // data-gradient operand [cin_pad][taps][cout_pad]
      q -= fwd_rows;
      const int ci = q / taps, t = q - ci * taps;
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(d.w_dgrad) + static_cast<size_t>(q) * d.cout_pad;
      const float* src = d.w_oihw + static_cast<size_t>(ci) * taps + t;
      for (int co = lane; co < d.cout_pad; co += 32)
        dst[co] = __float2bfloat16_rn((co < d.cout && ci < d.cin) ? src[static_cast<size_t>(co) * d.cin * taps] : 0.f);
    }
  }
}

extern "C" int yb200_pack_conv_weights_batched(const yb200_pack_desc* table_dev, const int64_t* prefix_dev, int n, int64_t total, void* stream) {
  YB_REQUIRE(table_dev && prefix_dev && n > 0 && n <= kPackMaxLayers && total > 0, YB200_ERR_INVALID, "pack_conv_weights_batched: bad arguments (n=%d)", n);
  const int blocks = static_cast<int>(std::min<long long>((2 * total + 255) / 256, 16LL * sm_count()));
  launch_k(pack_conv_weights_batched_kernel, blocks, 256, 0, as_stream(stream), table_dev, reinterpret_cast<const long long*>(prefix_dev), n);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int yb200_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, int cin_pad, void* w_fwd,
                                      void* w_dgrad, void* stream) {
  return pack_weight_impl(w_oihw, nullptr, cout, cin, ksize, cout_pad, cin_pad, w_fwd, w_dgrad, stream);
}

__global__ void scale_bias_kernel(const float* __restrict__ scale, const float* __restrict__ bias, int n, float* __restrict__ out) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = scale[i] * bias[i];
}

extern "C" int yb200_pack_conv_weight_scaled(const float* w_oihw, const float* cout_scale, const float* bias, int cout, int cin, int ksize,
                                             int cout_pad, int cin_pad, void* w_fwd, void* w_dgrad, float* scaled_bias, void* stream) {
  YB_REQUIRE(cout_scale != nullptr, YB200_ERR_INVALID, "pack_conv_weight_scaled: null scale");
  YB_REQUIRE((bias == nullptr) == (scaled_bias == nullptr), YB200_ERR_INVALID, "pack_conv_weight_scaled: pass both bias and scaled_bias or neither");
  if (bias) launch_k(scale_bias_kernel, ceil_div(cout, 256), 256, 0, as_stream(stream), cout_scale, bias, cout, scaled_bias);
  return pack_weight_impl(w_oihw, cout_scale, cout, cin, ksize, cout_pad, cin_pad, w_fwd, w_dgrad, stream);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// lo_delta > 0 selects the STRICT (operand-split) form: activations and weights are sums of `planes` bf16 values (x = x0 + x1 [+ x2], 8 more
// significant bits per plane: 16 bits with two planes, the full 24 of fp32 with three); plane j of an activation lives j * lo_delta channels
// after plane 0 in the same NHWC buffer and the weight matrix is [rows][plane 0 taps | plane 1 taps | plane 2 taps].  The product keeps every
// term x_i * w_j with i + j < planes (the dropped ones are below 2^-8planes relative) as extra taps of the SAME implicit GEMM -- one fp32
// accumulator in TMEM, no extra kernel: 3 taps per spatial tap for two planes, 6 for three.
static int conv_fwd_common(const yb200_act* x, const void* w_fwd, int cout, int ksize, int stride, ConvGemmParams& p,
                           cudaStream_t st, int lo_delta = 0, int planes = 1, const yb200_act* out_act = nullptr, int group = 0) {
  YB_REQUIRE(w_fwd != nullptr, YB200_ERR_INVALID, "conv fwd: null weights");
  YB_REQUIRE((ksize == 1 && stride == 1) || (ksize == 2 && stride == 2) || (ksize == 3 && (stride == 1 || stride == 2)), YB200_ERR_UNSUPPORTED,
             "conv fwd: ksize=%d stride=%d not implemented", ksize, stride);
  const int bk = pick_block_k(x->c);
  YB_REQUIRE(bk != 0, YB200_ERR_UNSUPPORTED, "conv fwd: input channels %d must be a multiple of 16", x->c);
  const int bn = pick_block_n(cout);
  const int oh = x->h / stride, ow = x->w / stride;
  set_tiles(p, x->n, oh, ow);
  p.num_taps = fill_fwd_taps(p.taps, *x, ksize, stride, x->c);
  long long kcols = 1LL * p.num_taps * x->c;
  if (lo_delta > 0) {
    YB_REQUIRE(planes == 2 || planes == 3, YB200_ERR_INVALID, "conv fwd (split): %d planes", planes);
    YB_REQUIRE(lo_delta % 8 == 0 && x->c_off + (planes - 1) * lo_delta + x->c <= x->c_pitch, YB200_ERR_INVALID,
               "conv fwd (split): plane %d [%d, %d) outside the channel pitch %d", planes - 1, x->c_off + (planes - 1) * lo_delta,
               x->c_off + (planes - 1) * lo_delta + x->c, x->c_pitch);
    const int nt = p.num_taps;
    const int plane_kb = nt * x->c;
    int terms[6][2], nterm = 0;  // (activation plane, weight plane), largest products first
    for (int sum = 0; sum < planes; ++sum)
      for (int i = 0; i <= sum; ++i) { terms[nterm][0] = i; terms[nterm][1] = sum - i; ++nterm; }
    for (int t = nt - 1; t >= 0; --t) {
      const ConvTap b = p.taps[t];
      for (int q = 0; q < nterm; ++q) {
        ConvTap e = b;
        e.c0 = b.c0 + terms[q][0] * lo_delta;
        e.kb = b.kb + terms[q][1] * plane_kb;
        p.taps[nterm * t + q] = e;
      }
    }
    p.num_taps = nterm * nt;
    kcols *= planes;
  }
  p.cin_blocks = x->c / bk;
  p.cout = cout;
  if (group > 1 && ksize == 3 && stride == 1 && lo_delta == 0 && p.cin_blocks == 1 && x->c_off == 0 && x->c == x->c_pitch && (x->c / group) % 16 == 0) {
    // pixel-grouped 3x3 convolution (yb200_conv2d_fwd_fold): of the left neighbour GROUP only its last pixel reaches this group's outputs, of the
    // right neighbour only its first.  The side taps therefore load / multiply `cpp` channels instead of group * cpp: their activation box starts
    // at the needed pixel (the rest of the box lies beyond the channel extent: TMA zero fill, no L2 traffic), the weight box at the matching
    // columns, and the persistent kernel issues only the first cpp / 16 K steps.  Correct in every kernel variant (the skipped products are
    // zero either way); YB200_STEM_SPARSE=0 keeps the dense taps for A/B runs.
    static int sparse = -1;
    if (sparse < 0) {
      const char* e = getenv("YB200_STEM_SPARSE");
      sparse = (e && e[0] == '0') ? 0 : 1;
    }
    const int cpp = x->c / group;
    for (int t = 0; t < p.num_taps && sparse; ++t) {
      ConvTap& tp = p.taps[t];
      if (tp.dw == -1) { tp.c0 += (group - 1) * cpp; tp.kb += (group - 1) * cpp; tp.ks = cpp / 16; }
      if (tp.dw == 1) tp.ks = cpp / 16;
    }
  }
  const int tw = 1 << p.log_tw, th = 1 << p.log_th, tn = 128 >> (p.log_tw + p.log_th);
  CUtensorMap tmA, tmB;
  int rc = make_act_map(&tmA, *x, stride == 2, bk, tw, th, tn);
  if (rc) return rc;
  // the CTA-pair kernel (column tile 256) loads the weight tile as two 128-row halves, one per CTA
  const bool pair = bn == 256 && use_pair_kernel() && !use_v1_kernel() && p.epi_mode != EPI_F32_BIAS;
  rc = make_mat_map(&tmB, w_fwd, cout, kcols, pair ? 128 : bn, bk);
  if (rc) return rc;
  const int num_kb = p.num_taps * p.cin_blocks;
  dim3 grid(p.tiles_w * p.tiles_h * p.tiles_n, ceil_div(cout, bn));
  return launch_conv(bn, bk, tmA, tmB, p, grid, pick_stages(bn, bk, num_kb), st, out_act);
}

static int conv2d_fwd_impl(const yb200_act* x, const void* w_fwd, const yb200_act* z, int ksize, int stride, double* stat_sum, double* stat_sqsum,
                           int stat_fold, void* stream);

extern "C" int yb200_conv2d_fwd(const yb200_act* x, const void* w_fwd, const yb200_act* z, int ksize, int stride,
                                double* stat_sum, double* stat_sqsum, void* stream) {
  return conv2d_fwd_impl(x, w_fwd, z, ksize, stride, stat_sum, stat_sqsum, 0, stream);
}

extern "C" int yb200_conv2d_fwd_fold(const yb200_act* x, const void* w_fwd, const yb200_act* z, int ksize, int stride, double* stat_sum,
                                     double* stat_sqsum, int stat_fold, void* stream) {
  YB_REQUIRE(stat_fold > 0 && z && z->c % stat_fold == 0, YB200_ERR_INVALID, "conv2d_fwd_fold: output channels must be a multiple of stat_fold");
  return conv2d_fwd_impl(x, w_fwd, z, ksize, stride, stat_sum, stat_sqsum, stat_fold, stream);
}

static int conv2d_fwd_impl(const yb200_act* x, const void* w_fwd, const yb200_act* z, int ksize, int stride, double* stat_sum, double* stat_sqsum,
                           int stat_fold, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv2d_fwd x"))) return rc;
  if ((rc = check_act(z, "conv2d_fwd z"))) return rc;
  YB_REQUIRE(stride == 1 || stride == 2, YB200_ERR_UNSUPPORTED, "conv2d_fwd: stride %d", stride);
  YB_REQUIRE(z->n == x->n && z->h * stride == x->h && z->w * stride == x->w, YB200_ERR_INVALID,
             "conv2d_fwd: output %dx%dx%d does not match input %dx%dx%d / stride %d", z->n, z->h, z->w, x->n, x->h, x->w, stride);
  YB_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), YB200_ERR_INVALID, "conv2d_fwd: pass both or neither stat buffer");
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  set_out_view(p, *z);
  p.epi_mode = stat_sum ? EPI_F16_STATS : EPI_F16;
  p.stat_sum = stat_sum;
  p.stat_sq = stat_sqsum;
  p.stat_fold = stat_fold;
  return conv_fwd_common(x, w_fwd, z->c, ksize, stride, p, as_stream(stream), 0, 1, z, stat_fold > 0 ? z->c / stat_fold : 0);
}

extern "C" int yb200_conv2d_bn_silu_fwd(const yb200_act* x, const void* w_fwd, const float* scale, const float* shift,
                                        const yb200_act* residual, const yb200_act* out, int ksize, int stride, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv2d_bn_silu_fwd x"))) return rc;
  if ((rc = check_act(out, "conv2d_bn_silu_fwd out"))) return rc;
  if (residual && (rc = check_act(residual, "conv2d_bn_silu_fwd residual"))) return rc;
  YB_REQUIRE(scale && shift, YB200_ERR_INVALID, "conv2d_bn_silu_fwd: null scale / shift");
  YB_REQUIRE(stride == 1 || stride == 2, YB200_ERR_UNSUPPORTED, "conv2d_bn_silu_fwd: stride %d", stride);
  YB_REQUIRE(out->n == x->n && out->h * stride == x->h && out->w * stride == x->w, YB200_ERR_INVALID,
             "conv2d_bn_silu_fwd: output %dx%dx%d does not match input %dx%dx%d / stride %d", out->n, out->h, out->w, x->n, x->h, x->w, stride);
  YB_REQUIRE(!residual || (residual->n == out->n && residual->h == out->h && residual->w == out->w && residual->c == out->c), YB200_ERR_INVALID,
             "conv2d_bn_silu_fwd: residual shape mismatch");
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  set_out_view(p, *out);
  p.epi_mode = EPI_BF16_BN_SILU;
  p.scale = scale;
  p.shift = shift;
  if (residual) {
    p.addend = static_cast<const __nv_bfloat16*>(residual->ptr) + residual->c_off;
    p.add_sw = residual->c_pitch;
    p.add_sh = 1LL * residual->c_pitch * residual->w;
    p.add_sn = 1LL * residual->c_pitch * residual->w * residual->h;
  }
  return conv_fwd_common(x, w_fwd, out->c, ksize, stride, p, as_stream(stream));
}

static void set_addend(ConvGemmParams& p, const yb200_act* a) {
  p.addend = static_cast<const __nv_bfloat16*>(a->ptr) + a->c_off;
  p.add_sw = a->c_pitch;
  p.add_sh = 1LL * a->c_pitch * a->w;
  p.add_sn = 1LL * a->c_pitch * a->w * a->h;
}
static bool same_geometry(const yb200_act* a, const yb200_act* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c && a->c_pitch == b->c_pitch;
}

extern "C" int yb200_conv2d_affine_fwd(const yb200_act* x, const void* w_fwd, const float* scale, const float* shift,
                                       const yb200_act* residual, const yb200_act* out, int ksize, int stride, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv2d_affine_fwd x"))) return rc;
  if ((rc = check_act(out, "conv2d_affine_fwd out"))) return rc;
  if (residual && (rc = check_act(residual, "conv2d_affine_fwd residual"))) return rc;
  YB_REQUIRE(stride == 1 || stride == 2, YB200_ERR_UNSUPPORTED, "conv2d_affine_fwd: stride %d", stride);
  YB_REQUIRE(out->n == x->n && out->h * stride == x->h && out->w * stride == x->w, YB200_ERR_INVALID,
             "conv2d_affine_fwd: output %dx%dx%d does not match input %dx%dx%d / stride %d", out->n, out->h, out->w, x->n, x->h, x->w, stride);
  YB_REQUIRE(!residual || (residual->n == out->n && residual->h == out->h && residual->w == out->w && residual->c == out->c), YB200_ERR_INVALID,
             "conv2d_affine_fwd: residual shape mismatch");
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  set_out_view(p, *out);
  p.epi_mode = EPI_BF16_AFFINE;
  p.scale = scale;
  p.shift = shift;
  if (residual) set_addend(p, residual);
  return conv_fwd_common(x, w_fwd, out->c, ksize, stride, p, as_stream(stream));
}

extern "C" int yb200_conv2d_relu_fwd(const yb200_act* x, const void* w_fwd, const float* bias, const yb200_act* out, int ksize, int stride,
                                     void* stream) {
  int rc;
  if ((rc = check_act(x, "conv2d_relu_fwd x"))) return rc;
  if ((rc = check_act(out, "conv2d_relu_fwd out"))) return rc;
  YB_REQUIRE(stride == 1 || stride == 2, YB200_ERR_UNSUPPORTED, "conv2d_relu_fwd: stride %d", stride);
  YB_REQUIRE(out->n == x->n && out->h * stride == x->h && out->w * stride == x->w, YB200_ERR_INVALID,
             "conv2d_relu_fwd: output %dx%dx%d does not match input %dx%dx%d / stride %d", out->n, out->h, out->w, x->n, x->h, x->w, stride);
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  set_out_view(p, *out);
  p.epi_mode = EPI_BF16_BIAS_RELU;
  p.shift = bias;
  return conv_fwd_common(x, w_fwd, out->c, ksize, stride, p, as_stream(stream));
}

extern "C" int yb200_linear_relu_fwd(const yb200_act* x, const void* w_fwd, const float* bias, const yb200_act* h_out, void* stream) {
  return yb200_conv2d_relu_fwd(x, w_fwd, bias, h_out, 1, 1, stream);
}

extern "C" int yb200_conv1x1_nchw_f32(const yb200_act* x, const void* w_fwd, const float* bias, int cout, float* out_nchw, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv1x1_nchw_f32 x"))) return rc;
  YB_REQUIRE(out_nchw && cout > 0 && cout <= 128, YB200_ERR_INVALID, "conv1x1_nchw_f32: bad arguments (cout=%d)", cout);
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  const long long hw = 1LL * x->h * x->w;
  p.out = out_nchw;                 // element (n, y, x, c) at n*cout*hw + c*hw + y*w + x: lanes (pixels) write consecutive floats per channel
  p.out_sn = 1LL * cout * hw;
  p.out_sh = x->w;
  p.out_sw = 1;
  YB_REQUIRE(hw < (1LL << 31), YB200_ERR_UNSUPPORTED, "conv1x1_nchw_f32: plane too large");
  p.out_sc = static_cast<int>(hw);
  p.out_mh = 1; p.out_mw = 1;
  p.bias = bias;                    // may be null (staged as zeros)
  p.epi_mode = EPI_F32_BIAS;
  return conv_fwd_common(x, w_fwd, cout, 1, 1, p, as_stream(stream));
}

// the same with one weight matrix PER IMAGE (w_fwd: [n][cout][cin] bf16, i.e. n * cout rows): torch.bmm(pred_kernel, mask_features) of a whole batch
// in one launch.  Pixel tiles must not span images (h * w a multiple of the 128-pixel tile: choose_tile then keeps tiles inside one image).
extern "C" int yb200_conv1x1_nchw_f32_batched(const yb200_act* x, const void* w_fwd, int cout, float* out_nchw, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv1x1_nchw_f32_batched x"))) return rc;
  YB_REQUIRE(out_nchw && cout > 0 && cout <= 128, YB200_ERR_INVALID, "conv1x1_nchw_f32_batched: bad arguments (cout=%d)", cout);
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  const long long hw = 1LL * x->h * x->w;
  p.out = out_nchw;
  p.out_sn = 1LL * cout * hw;
  p.out_sh = x->w;
  p.out_sw = 1;
  YB_REQUIRE(hw < (1LL << 31), YB200_ERR_UNSUPPORTED, "conv1x1_nchw_f32_batched: plane too large");
  p.out_sc = static_cast<int>(hw);
  p.out_mh = 1; p.out_mw = 1;
  p.epi_mode = EPI_F32_BIAS;
  p.b_img_rows = cout;
  set_tiles(p, x->n, x->h, x->w);
  YB_REQUIRE(p.log_tw + p.log_th == 7, YB200_ERR_UNSUPPORTED,
             "conv1x1_nchw_f32_batched: a 128-pixel tile would span images at %dx%d (use yb200_conv1x1_nchw_f32 per image)", x->h, x->w);
  YB_REQUIRE(pick_block_n(cout) <= 128 && !use_v1_kernel(), YB200_ERR_UNSUPPORTED, "conv1x1_nchw_f32_batched: needs the persistent kernel");
  // conv_fwd_common, with the weight matrix map covering all images' rows
  const int bk = pick_block_k(x->c);
  YB_REQUIRE(bk != 0, YB200_ERR_UNSUPPORTED, "conv1x1_nchw_f32_batched: input channels %d must be a multiple of 16", x->c);
  const int bn = pick_block_n(cout);
  p.num_taps = fill_fwd_taps(p.taps, *x, 1, 1, x->c);
  p.cin_blocks = x->c / bk;
  p.cout = cout;
  const int tw = 1 << p.log_tw, th = 1 << p.log_th;
  CUtensorMap tmA, tmB;
  if ((rc = make_act_map(&tmA, *x, false, bk, tw, th, 1))) return rc;
  if ((rc = make_mat_map(&tmB, w_fwd, 1LL * x->n * cout, x->c, bn, bk))) return rc;
  dim3 grid(p.tiles_w * p.tiles_h * p.tiles_n, ceil_div(cout, bn));
  return launch_conv(bn, bk, tmA, tmB, p, grid, pick_stages(bn, bk, p.cin_blocks), as_stream(stream));
}

extern "C" int yb200_linear_gelu_fwd(const yb200_act* x, const void* w_fwd, const float* bias, const yb200_act* u_out, const yb200_act* h_out,
                                     void* stream) {
  int rc;
  if ((rc = check_act(x, "linear_gelu_fwd x"))) return rc;
  if ((rc = check_act(h_out, "linear_gelu_fwd h"))) return rc;
  if (u_out && (rc = check_act(u_out, "linear_gelu_fwd u"))) return rc;
  YB_REQUIRE(h_out->n == x->n && h_out->h == x->h && h_out->w == x->w, YB200_ERR_INVALID, "linear_gelu_fwd: pixel grids differ");
  YB_REQUIRE(!u_out || same_geometry(u_out, h_out), YB200_ERR_INVALID, "linear_gelu_fwd: u and h must have the same shape and channel pitch");
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  set_out_view(p, *h_out);
  p.epi_mode = EPI_BF16_BIAS_GELU;
  p.shift = bias;
  if (u_out) p.aux_out = static_cast<__nv_bfloat16*>(u_out->ptr) + u_out->c_off;
  return conv_fwd_common(x, w_fwd, h_out->c, 1, 1, p, as_stream(stream));
}

extern "C" int yb200_conv1x1_bias_f32(const yb200_act* x, const void* w_fwd, const float* bias, int cout, float* out,
                                      int a_total, int a_off, int c_total, int c_off, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv1x1_bias_f32 x"))) return rc;
  YB_REQUIRE(bias && out && cout > 0 && cout <= 128, YB200_ERR_INVALID, "conv1x1_bias_f32: bad arguments (cout=%d)", cout);
  YB_REQUIRE(a_off >= 0 && a_off + x->h * x->w <= a_total && c_off >= 0 && c_off + cout <= c_total, YB200_ERR_INVALID,
             "conv1x1_bias_f32: slice [%d+%d, %d+%d] outside [%d, %d]", a_off, x->h * x->w, c_off, cout, a_total, c_total);
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = out + 1LL * a_off * c_total + c_off;
  p.out_sn = 1LL * a_total * c_total;
  p.out_sh = 1LL * x->w * c_total;
  p.out_sw = c_total;
  p.out_sc = 1;
  p.out_mh = 1; p.out_mw = 1;
  p.bias = bias;
  p.epi_mode = EPI_F32_BIAS;
  return conv_fwd_common(x, w_fwd, cout, 1, 1, p, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// strict (operand-split) forward: fp32 results from bf16 tensor-core products
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_split_kernel(const float* __restrict__ w, int cout, int cin, int taps, int cout_pad, int cin_pad, int planes,
                                              __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const long long per_plane = 1LL * taps * cin_pad;
  const long long total = 1LL * cout_pad * per_plane;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = static_cast<int>(i % cin_pad);
    const int t = static_cast<int>((i / cin_pad) % taps);
    const int co = static_cast<int>(i / per_plane);
    float r = (co < cout && ci < cin) ? w[(1LL * co * cin + ci) * taps + t] : 0.f;
    __nv_bfloat16* row = out + planes * co * per_plane + 1LL * t * cin_pad + ci;
    for (int pl = 0; pl < planes; ++pl) {
      const __nv_bfloat16 h = __float2bfloat16_rn(r);
      row[pl * per_plane] = h;
      r -= __bfloat162float(h);  // exact: the residual of a bf16 rounding is representable in fp32
    }
  }
}

extern "C" int yb200_pack_conv_weight_split(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, int cin_pad, int planes, void* w_split,
                                            void* stream) {
  YB_REQUIRE(w_oihw && w_split, YB200_ERR_INVALID, "pack_conv_weight_split: null pointer");
  YB_REQUIRE(cout > 0 && cin > 0 && (ksize >= 1 && ksize <= 3) && cout_pad >= cout && cin_pad >= cin && (planes == 2 || planes == 3), YB200_ERR_INVALID,
             "pack_conv_weight_split: bad sizes cout=%d cin=%d k=%d pads=%d,%d planes=%d", cout, cin, ksize, cout_pad, cin_pad, planes);
  const long long total = 1LL * cout_pad * ksize * ksize * cin_pad;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 4096));
  launch_k(pack_conv_weight_split_kernel, blocks, 256, 0, as_stream(stream), w_oihw, cout, cin, ksize * ksize, cout_pad, cin_pad, planes,
                                                                       static_cast<__nv_bfloat16*>(w_split));
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int yb200_conv2d_fwd_split(const yb200_act* x, int lo_delta, int planes, const void* w_split, int cout, int ksize, int stride, float* z,
                                      int z_pitch, int z_off, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv2d_fwd_split x"))) return rc;
  YB_REQUIRE(z && cout > 0 && z_off >= 0 && z_off + cout <= z_pitch && lo_delta > 0, YB200_ERR_INVALID,
             "conv2d_fwd_split: bad output slice [%d, %d) of %d (lo_delta %d)", z_off, z_off + cout, z_pitch, lo_delta);
  YB_REQUIRE(stride == 1 || stride == 2, YB200_ERR_UNSUPPORTED, "conv2d_fwd_split: stride %d", stride);
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  const int oh = x->h / stride, ow = x->w / stride;
  p.out = z + z_off;  // fp32 NHWC with a channel pitch
  p.out_sw = z_pitch;
  p.out_sh = 1LL * z_pitch * ow;
  p.out_sn = 1LL * z_pitch * ow * oh;
  p.out_sc = 1;
  p.out_mh = 1; p.out_mw = 1;
  p.epi_mode = EPI_F32_BIAS;  // bias == null: staged as zeros
  return conv_fwd_common(x, w_split, cout, ksize, stride, p, as_stream(stream), lo_delta, planes);
}

extern "C" int yb200_conv1x1_bias_f32_split(const yb200_act* x, int lo_delta, int planes, const void* w_split, const float* bias, int cout, float* out,
                                            int a_total, int a_off, int c_total, int c_off, void* stream) {
  int rc;
  if ((rc = check_act(x, "conv1x1_bias_f32_split x"))) return rc;
  YB_REQUIRE(bias && out && cout > 0 && cout <= 128 && lo_delta > 0, YB200_ERR_INVALID, "conv1x1_bias_f32_split: bad arguments (cout=%d)", cout);
  YB_REQUIRE(a_off >= 0 && a_off + x->h * x->w <= a_total && c_off >= 0 && c_off + cout <= c_total, YB200_ERR_INVALID,
             "conv1x1_bias_f32_split: slice [%d+%d, %d+%d] outside [%d, %d]", a_off, x->h * x->w, c_off, cout, a_total, c_total);
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = out + 1LL * a_off * c_total + c_off;
  p.out_sn = 1LL * a_total * c_total;
  p.out_sh = 1LL * x->w * c_total;
  p.out_sw = c_total;
  p.out_sc = 1;
  p.out_mh = 1; p.out_mw = 1;
  p.bias = bias;
  p.epi_mode = EPI_F32_BIAS;
  return conv_fwd_common(x, w_split, cout, 1, 1, p, as_stream(stream), lo_delta, planes);
}

// ------------------------------------------------------------------------------------------------
// data gradient
// ------------------------------------------------------------------------------------------------
static int dgrad_impl(const yb200_act* dz, const void* w_dgrad, const yb200_act* dx, const yb200_act* addend, int ksize, int stride,
                      const yb200_act* gelu_u, double* colsum, void* stream, int act_mode = EPI_BF16_GELU_BWD, int num_seg = 0,
                      const yb200_bnbwd_seg* segs = nullptr) {
  int rc;
  if ((rc = check_act(dz, "conv2d_dgrad dz"))) return rc;
  if ((rc = check_act(dx, "conv2d_dgrad dx"))) return rc;
  if (addend && (rc = check_act(addend, "conv2d_dgrad addend"))) return rc;
  YB_REQUIRE(w_dgrad != nullptr, YB200_ERR_INVALID, "conv2d_dgrad: null weights");
  YB_REQUIRE((ksize == 1 && stride == 1) || (ksize == 2 && stride == 2) || (ksize == 3 && (stride == 1 || stride == 2)), YB200_ERR_UNSUPPORTED,
             "conv2d_dgrad: ksize=%d stride=%d not implemented", ksize, stride);
  YB_REQUIRE(dz->n == dx->n && dz->h * stride == dx->h && dz->w * stride == dx->w, YB200_ERR_INVALID,
             "conv2d_dgrad: dz %dx%dx%d vs dx %dx%dx%d stride %d", dz->n, dz->h, dz->w, dx->n, dx->h, dx->w, stride);
  YB_REQUIRE(!addend || (addend->n == dx->n && addend->h == dx->h && addend->w == dx->w && addend->c == dx->c), YB200_ERR_INVALID,
             "conv2d_dgrad: addend shape mismatch");
  const int bk = pick_block_k(dz->c);
  YB_REQUIRE(bk != 0, YB200_ERR_UNSUPPORTED, "conv2d_dgrad: dz channels %d must be a multiple of 16", dz->c);
  const int cin = dx->c;
  const int bn = pick_block_n(cin);
  const int taps_total = ksize * ksize;
  cudaStream_t st = as_stream(stream);

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  set_out_view(p, *dx);
  p.epi_mode = EPI_BF16;
  if (gelu_u) {
    p.epi_mode = act_mode;
    p.aux_in = static_cast<const __nv_bfloat16*>(gelu_u->ptr) + gelu_u->c_off;
    p.stat_sum = colsum;
  }
  p.cout = cin;
  p.cin_blocks = dz->c / bk;
  if (addend) {
    p.addend = static_cast<const __nv_bfloat16*>(addend->ptr) + addend->c_off;
    p.add_sw = addend->c_pitch;
    p.add_sh = 1LL * addend->c_pitch * addend->w;
    p.add_sn = 1LL * addend->c_pitch * addend->w * addend->h;
  }
  if (num_seg > 0) {
    YB_REQUIRE(num_seg <= 2 && segs != nullptr && gelu_u == nullptr, YB200_ERR_INVALID, "conv2d_dgrad_bnbwd: 1 or 2 segments");
    YB_REQUIRE(bn <= 128, YB200_ERR_UNSUPPORTED, "conv2d_dgrad_bnbwd: gradient tensors of %d channels use a 256-wide column tile (not supported)", cin);
    p.num_bnseg = num_seg;
    for (int i = 0; i < num_seg; ++i) {
      const yb200_bnbwd_seg& sgm = segs[i];
      const yb200_act& z = sgm.z;
      YB_REQUIRE(z.ptr && sgm.scale && sgm.shift && sgm.sum_du && sgm.sum_duz, YB200_ERR_INVALID, "conv2d_dgrad_bnbwd: null pointer in segment %d", i);
      YB_REQUIRE(z.n == dx->n && z.h == dx->h && z.w == dx->w && z.c % 32 == 0 && sgm.dx_c_begin % 32 == 0 && sgm.dx_c_begin >= 0 &&
                     sgm.dx_c_begin + z.c <= dx->c && z.c_off % 8 == 0 && z.c_pitch % 8 == 0,
                 YB200_ERR_INVALID, "conv2d_dgrad_bnbwd: segment %d (z %dx%dx%dx%d at dx channel %d) does not fit dx %dx%dx%dx%d", i, z.n, z.h, z.w, z.c,
                 sgm.dx_c_begin, dx->n, dx->h, dx->w, dx->c);
      BnBwdSeg& d = p.bnseg[i];
      d.col_begin = sgm.dx_c_begin;
      d.col_end = sgm.dx_c_begin + z.c;
      d.z = static_cast<const __half*>(z.ptr) + z.c_off;
      d.z_sw = z.c_pitch;
      d.z_sh = 1LL * z.c_pitch * z.w;
      d.z_sn = 1LL * z.c_pitch * z.w * z.h;
      d.scale = sgm.scale; d.shift = sgm.shift; d.sum_du = sgm.sum_du; d.sum_duz = sgm.sum_duz;
    }
    YB_REQUIRE(num_seg < 2 || p.bnseg[0].col_end <= p.bnseg[1].col_begin || p.bnseg[1].col_end <= p.bnseg[0].col_begin, YB200_ERR_INVALID,
               "conv2d_dgrad_bnbwd: overlapping segments");
  }
  set_tiles(p, dz->n, dz->h, dz->w);  // pixel grid = dz grid (for stride 2: one output-parity class at a time)
  const int tw = 1 << p.log_tw, th = 1 << p.log_th, tn = 128 >> (p.log_tw + p.log_th);
  CUtensorMap tmA, tmB;
  if ((rc = make_act_map(&tmA, *dz, false, bk, tw, th, tn))) return rc;
  if ((rc = make_mat_map(&tmB, w_dgrad, cin, 1LL * taps_total * dz->c, (bn == 256 && use_pair_kernel() && !use_v1_kernel() && num_seg == 0) ? 128 : bn, bk))) return rc;
  dim3 grid(p.tiles_w * p.tiles_h * p.tiles_n, ceil_div(cin, bn));

  if (stride == 1) {
    int nt = 0;
    if (ksize == 1) {
      p.taps[nt++] = ConvTap{dz->c_off, 0, 0, 0, 0};
    } else {
      // dx[y,x] = sum_k dz[y + 1 - kh, x + 1 - kw] * W[kh,kw]
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) p.taps[nt++] = ConvTap{dz->c_off, 1 - kw, 0, 1 - kh, (kh * 3 + kw) * dz->c};
    }
    p.num_taps = nt;
    return launch_conv(bn, bk, tmA, tmB, p, grid, pick_stages(bn, bk, nt * p.cin_blocks), st, (gelu_u == nullptr && addend == nullptr) ? dx : nullptr);
  }
  // stride 2: input pixel (2i+ph, 2j+pw) receives  kh with (ph + 1 - kh) even:  ph=0 -> kh=1 (row i);  ph=1 -> kh=0 (row i+1), kh=2 (row i)
  auto phase_taps = [&](int ph, int pw, ConvTap* out) {
    int nt = 0;
    if (ksize == 2) out[nt++] = ConvTap{dz->c_off, 0, 0, 0, (ph * 2 + pw) * dz->c};  // 2x2 s2: input pixel (2i+ph, 2j+pw) sees only tap (ph, pw)
    for (int kh = 0; kh < 3 && ksize == 3; ++kh) {
      if (((ph + 1 - kh) & 1) != 0) continue;
      const int dh = (ph + 1 - kh) / 2;
      for (int kw = 0; kw < 3; ++kw) {
        if (((pw + 1 - kw) & 1) != 0) continue;
        const int dw = (pw + 1 - kw) / 2;
        out[nt++] = ConvTap{dz->c_off, dw, 0, dh, (kh * 3 + kw) * dz->c};
      }
    }
    return nt;
  };
  p.out_mh = 2; p.out_mw = 2;
  static int one_launch = -1;  // YB200_DGRAD_PHASES=4 restores one launch per output-parity class (A/B runs)
  if (one_launch < 0) {
    const char* e = getenv("YB200_DGRAD_PHASES");
    one_launch = (e && e[0] == '4') ? 0 : 1;
  }
  const bool pair = bn == 256 && use_pair_kernel() && num_seg == 0;  // the CTA-pair kernel keeps the per-phase launches
  if (one_launch && !use_v1_kernel() && !pair) {
    // all four phases in ONE persistent launch: the phases of a pixel tile run back to back on neighbouring CTAs and share its dz tile in L2
    int nt = 0;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        p.phase_tap[ph * 2 + pw] = nt;
        nt += phase_taps(ph, pw, p.taps + nt);
      }
    p.phase_tap[4] = nt;
    p.num_taps = nt;
    p.num_phases = 4;
    return launch_conv(bn, bk, tmA, tmB, p, grid, pick_stages(bn, bk, p.cin_blocks), st);
  }
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int nt = phase_taps(ph, pw, p.taps);
      p.num_taps = nt;
      p.out_ph = ph; p.out_pw = pw;
      if ((rc = launch_conv(bn, bk, tmA, tmB, p, grid, pick_stages(bn, bk, nt * p.cin_blocks), st))) return rc;
    }
  return 0;
}

extern "C" int yb200_conv2d_dgrad(const yb200_act* dz, const void* w_dgrad, const yb200_act* dx, const yb200_act* addend,
                                  int ksize, int stride, void* stream) {
  return dgrad_impl(dz, w_dgrad, dx, addend, ksize, stride, nullptr, nullptr, stream);
}

extern "C" int yb200_conv2d_dgrad_bnbwd(const yb200_act* dz, const void* w_dgrad, const yb200_act* dx, const yb200_act* addend, int ksize, int stride,
                                        int num_segments, const yb200_bnbwd_seg* segments, void* stream) {
  YB_REQUIRE(num_segments >= 1, YB200_ERR_INVALID, "conv2d_dgrad_bnbwd: no segments (use yb200_conv2d_dgrad)");
  return dgrad_impl(dz, w_dgrad, dx, addend, ksize, stride, nullptr, nullptr, stream, EPI_BF16, num_segments, segments);
}

extern "C" int yb200_linear_dgrad_gelu(const yb200_act* dh_src, const void* w_dgrad, const yb200_act* u, const yb200_act* du, double* bias_grad_sum,
                                       void* stream) {
  int rc;
  if ((rc = check_act(u, "linear_dgrad_gelu u"))) return rc;
  if ((rc = check_act(du, "linear_dgrad_gelu du"))) return rc;
  YB_REQUIRE(same_geometry(u, du), YB200_ERR_INVALID, "linear_dgrad_gelu: u and du must have the same shape and channel pitch");
  return dgrad_impl(dh_src, w_dgrad, du, nullptr, 1, 1, u, bias_grad_sum, stream);
}

extern "C" int yb200_linear_dgrad_relu(const yb200_act* dz, const void* w_dgrad, const yb200_act* h, const yb200_act* du, double* bias_grad_sum,
                                       void* stream) {
  int rc;
  if ((rc = check_act(h, "linear_dgrad_relu h"))) return rc;
  if ((rc = check_act(du, "linear_dgrad_relu du"))) return rc;
  YB_REQUIRE(same_geometry(h, du), YB200_ERR_INVALID, "linear_dgrad_relu: h and du must have the same shape and channel pitch");
  return dgrad_impl(dz, w_dgrad, du, nullptr, 1, 1, h, bias_grad_sum, stream, EPI_BF16_RELU_BWD);
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
namespace {
struct WgradPlan {
  WgradParams p;
  int tmem_cols;
  int splits;
  int smem;
  int tw, th, tn;
};

int plan_wgrad(const yb200_act* x, const yb200_act* dz, int ksize, int stride, WgradPlan* pl, int group = 0) {
  int rc;
  if ((rc = check_act(x, "conv2d_wgrad x"))) return rc;
  if ((rc = check_act(dz, "conv2d_wgrad dz"))) return rc;
  YB_REQUIRE((ksize == 1 && stride == 1) || (ksize == 2 && stride == 2) || (ksize == 3 && (stride == 1 || stride == 2)), YB200_ERR_UNSUPPORTED,
             "conv2d_wgrad: ksize=%d stride=%d not implemented", ksize, stride);
  YB_REQUIRE(dz->n == x->n && dz->h * stride == x->h && dz->w * stride == x->w, YB200_ERR_INVALID,
             "conv2d_wgrad: dz %dx%dx%d vs x %dx%dx%d stride %d", dz->n, dz->h, dz->w, x->n, x->h, x->w, stride);
  YB_REQUIRE(x->c % 16 == 0, YB200_ERR_UNSUPPORTED, "conv2d_wgrad: input channels %d must be a multiple of 16", x->c);
  YB_REQUIRE(dz->c % 16 == 0, YB200_ERR_UNSUPPORTED, "conv2d_wgrad: output channels %d must be a multiple of 16", dz->c);
  WgradParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.cout = dz->c;
  p.cin = x->c;
  p.kc_a = dz->c >= 64 ? 64 : (dz->c >= 32 ? 32 : 16);
  YB_REQUIRE(dz->c % p.kc_a == 0 || dz->c == dz->c_pitch, YB200_ERR_UNSUPPORTED,
             "conv2d_wgrad: dz channel slice %d not a multiple of %d", dz->c, p.kc_a);
  // boxes of kc_a channels per 128-row tile; a box that overshoots the slice (48 -> 2 x 32, 80 -> 2 x 64, 96 -> 2 x 64) is legal only when
  // the slice ends at the channel pitch (required above): the overshoot is then out of bounds for TMA and zero filled
  p.ma = ceil_div(dz->c < 128 ? dz->c : 128, p.kc_a);
  p.cout_tiles = ceil_div(dz->c, 128);
  p.kc_b = x->c % 64 == 0 ? 64 : (x->c % 32 == 0 ? 32 : 16);
  p.bn = x->c % 128 == 0 ? 128 : (x->c % 64 == 0 ? 64 : p.kc_b);
  p.nb = p.bn / p.kc_b;
  p.cin_tiles = x->c / p.bn;
  p.num_taps = fill_fwd_taps(p.taps, *x, ksize, stride, 0);
  if (group > 1) {
    // pixel-grouped 3x3 convolution (the stem): the expanded weight matrix is non-zero for the left / right neighbour group only at its last /
    // first pixel, and only those entries are folded back onto the parameter.  The side taps therefore load and multiply cpp = C / group channels.
    const int cpp = x->c / group;
    YB_REQUIRE(ksize == 3 && stride == 1 && x->c % group == 0 && cpp % 16 == 0 && x->c_off == 0 && x->c == x->c_pitch && p.cin_tiles == 1 && p.nb == 1,
               YB200_ERR_UNSUPPORTED, "conv2d_wgrad_grouped: needs a 3x3 stride-1 convolution on a whole [.., %d x 16k]-channel grouped tensor of <= 64 channels (got c=%d pitch=%d group=%d)",
               group, x->c, x->c_pitch, group);
    static int sparse = -1;
    if (sparse < 0) {
      const char* e = getenv("YB200_STEM_SPARSE");
      sparse = (e && e[0] == '0') ? 0 : 1;
    }
    for (int t = 0; t < p.num_taps && sparse; ++t) {
      ConvTap& tp = p.taps[t];
      tp.kb = 0;
      if (tp.dw == -1) { tp.c0 += (group - 1) * cpp; tp.kb = (group - 1) * cpp; tp.ks = cpp / 16; }
      if (tp.dw == 1) tp.ks = cpp / 16;
    }
  }
  p.tpc = p.num_taps == 1 ? 1 : 3;  // (9 taps per CTA for narrow layers was measured slower: many 2 KB TMA boxes per stage)
  p.tap_groups = ceil_div(p.num_taps, p.tpc);
  p.dz_c0 = dz->c_off;
  choose_tile(dz->n, dz->h, dz->w, kWgPix, &p.log_tw, &p.log_th);
  pl->tw = 1 << p.log_tw; pl->th = 1 << p.log_th; pl->tn = kWgPix >> (p.log_tw + p.log_th);
  p.tiles_w = ceil_div(dz->w, pl->tw); p.tiles_h = ceil_div(dz->h, pl->th); p.tiles_n = ceil_div(dz->n, pl->tn);
  p.num_blocks = p.tiles_w * p.tiles_h * p.tiles_n;
  const int base = p.cout_tiles * p.cin_tiles * p.tap_groups;
  int cols = p.tpc * p.bn;
  int tc = 32;
  while (tc < cols) tc <<= 1;
  pl->tmem_cols = tc;
  const int stage = kWgPix * 2 * (p.kc_a * p.ma + p.kc_b * p.nb * p.tpc);
  p.stages = kWgStages;
  if (std::min(512 / tc, (220 * 1024) / (stage * kWgStages + 1024)) <= 1) {  // one CTA per SM anyway: deepen the ring (YB200_WGRAD_STAGES caps it, A/B)
    static int cap = -1;
    if (cap < 0) {
      const char* e = getenv("YB200_WGRAD_STAGES");
      cap = e ? atoi(e) : kWgMaxStages;
      if (cap < kWgStages) cap = kWgStages;
      if (cap > kWgMaxStages) cap = kWgMaxStages;
    }
    p.stages = std::max(kWgStages, std::min(cap, (220 * 1024 - 1024) / stage));
  }
  pl->smem = stage * p.stages + 1024;
  // Split the pixel range so that ONE wave of CTAs covers the machine (every CTA pays TMEM allocation, pipeline fill and
  // a full accumulator write-back, so extra waves are pure overhead); occupancy is bounded by TMEM columns and shared memory.
  int occ = std::min(std::min(512 / tc, (220 * 1024) / pl->smem), 8);
  if (occ < 1) occ = 1;
  int splits = (occ * sm_count()) / base;  // floor: a partial second wave would double the tail
  if (splits > p.num_blocks / 8) splits = p.num_blocks / 8;  // at least 8 pixel blocks per CTA
  const long long wbytes = 4LL * p.cout * p.num_taps * p.cin;
  const long long max_splits = (128LL << 20) / wbytes;
  if (splits > max_splits) splits = static_cast<int>(max_splits);
  if (splits < 1) splits = 1;
  p.blocks_per_split = ceil_div(p.num_blocks, splits);
  pl->splits = ceil_div(p.num_blocks, p.blocks_per_split);
  return 0;
}
}  // namespace

extern "C" int64_t yb200_conv2d_wgrad_workspace(const yb200_act* x, const yb200_act* dz, int ksize, int stride) {
  WgradPlan pl;
  int rc = plan_wgrad(x, dz, ksize, stride, &pl);
  if (rc) return rc;
  return 4LL * pl.splits * pl.p.cout * pl.p.num_taps * pl.p.cin;
}

template <int TC>
static int launch_wgrad_inst(const CUtensorMap& tmDz, const CUtensorMap& tmX, const WgradPlan& pl, dim3 grid, cudaStream_t st) {
  static PerDevice<int> max_set_dev(0);
  int& max_set = max_set_dev.cur();
  if (pl.smem > max_set) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    max_set = pl.smem;
  }
  launch_k_opt(use_pdl_wgrad(), wgrad_gemm_kernel<TC>, grid, kConvThreads, pl.smem, st, tmDz, tmX, pl.p);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int wgrad_impl(const yb200_act* x, const yb200_act* dz, int ksize, int stride, int cin_real, int group, float* grad_oihw, int accumulate,
                      void* workspace, int64_t workspace_bytes, void* stream);
extern "C" int yb200_conv2d_wgrad(const yb200_act* x, const yb200_act* dz, int ksize, int stride, int cin_real, float* grad_oihw,
                                  int accumulate, void* workspace, int64_t workspace_bytes, void* stream) {
  return wgrad_impl(x, dz, ksize, stride, cin_real, 0, grad_oihw, accumulate, workspace, workspace_bytes, stream);
}
extern "C" int yb200_conv2d_wgrad_grouped(const yb200_act* x, const yb200_act* dz, int ksize, int stride, int cin_real, int group, float* grad_oihw,
                                          int accumulate, void* workspace, int64_t workspace_bytes, void* stream) {
  YB_REQUIRE(group > 1, YB200_ERR_INVALID, "conv2d_wgrad_grouped: group %d (use yb200_conv2d_wgrad)", group);
  return wgrad_impl(x, dz, ksize, stride, cin_real, group, grad_oihw, accumulate, workspace, workspace_bytes, stream);
}
static int wgrad_impl(const yb200_act* x, const yb200_act* dz, int ksize, int stride, int cin_real, int group, float* grad_oihw, int accumulate,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  WgradPlan pl;
  int rc = plan_wgrad(x, dz, ksize, stride, &pl, group);
  if (rc) return rc;
  YB_REQUIRE(grad_oihw && workspace, YB200_ERR_INVALID, "conv2d_wgrad: null pointer");
  YB_REQUIRE(cin_real > 0 && cin_real <= x->c, YB200_ERR_INVALID, "conv2d_wgrad: cin_real %d vs padded %d", cin_real, x->c);
  const int64_t need = 4LL * pl.splits * pl.p.cout * pl.p.num_taps * pl.p.cin;
  YB_REQUIRE(workspace_bytes >= need, YB200_ERR_INVALID, "conv2d_wgrad: workspace %lld < %lld bytes", (long long)workspace_bytes,
             (long long)need);
  pl.p.ws = static_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  CUtensorMap tmDz, tmX;
  if ((rc = make_act_map(&tmDz, *dz, false, pl.p.kc_a, pl.tw, pl.th, pl.tn))) return rc;
  if ((rc = make_act_map(&tmX, *x, stride == 2, pl.p.kc_b, pl.tw, pl.th, pl.tn))) return rc;
  dim3 grid(pl.p.cout_tiles * pl.p.cin_tiles * pl.p.tap_groups, pl.splits);
  switch (pl.tmem_cols) {
    case 32: rc = launch_wgrad_inst<32>(tmDz, tmX, pl, grid, st); break;
    case 64: rc = launch_wgrad_inst<64>(tmDz, tmX, pl, grid, st); break;
    case 128: rc = launch_wgrad_inst<128>(tmDz, tmX, pl, grid, st); break;
    case 256: rc = launch_wgrad_inst<256>(tmDz, tmX, pl, grid, st); break;
    case 512: rc = launch_wgrad_inst<512>(tmDz, tmX, pl, grid, st); break;
    default: return fail(YB200_ERR_UNSUPPORTED, "conv2d_wgrad: %d TMEM columns", pl.tmem_cols);
  }
  if (rc) return rc;
  const long long total = 1LL * pl.p.cout * pl.p.num_taps * pl.p.cin;
  const int blocks = static_cast<int>(std::min<long long>((total + 31) / 32, 16 * sm_count()));
  launch_k_opt(use_pdl_wgrad(), wgrad_reduce_kernel, blocks, 256, 0, st, pl.p.ws, grad_oihw, pl.splits, pl.p.cout, pl.p.num_taps, pl.p.cin, cin_real, accumulate);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Weight-gradient GEMM on tcgen05 (sm_100a).
//
//   dW_tap[co, ci] = sum over pixels  dz[px, co] * x_tap[px, ci]
//
// Both operands are NHWC bf16 tensors, i.e. the contraction index (pixel) is the *slow* index of each
// smem tile: UMMA "MN-major" descriptors.  dz tiles are dense 64-pixel boxes; x tiles are the same boxes
// displaced by the tap (identical TMA maps / tap tables as the forward kernel), so zero padding and stride-2
// need no extra code.  One CTA owns (cout tile of 128) x (cin tile of BN) x (tap group) x (pixel range) and keeps
// one fp32 accumulator per tap in TMEM; partial tiles go to a split-K workspace reduced by wgrad_reduce_kernel
// (deterministic, no atomics).
#pragma once
#include "sm100.cuh"
#include "conv_gemm.cuh"

namespace yb {

constexpr int kWgPix = 64;      // pixels per K block (4 UMMA K-steps)
constexpr int kWgMaxTpc = 9;    // taps per CTA (9 when the cin tile is narrow enough for 9 accumulators in TMEM, else 3)
constexpr int kWgStages = 3;     // default ring depth (keeps two or more CTAs per SM on the small layers)
constexpr int kWgMaxStages = 6;  // one-CTA-per-SM plans (the pixel-grouped stem: 40 KB stages) take as many as fit: the kernel is latency bound otherwise

struct WgradParams {
  int tiles_w, tiles_h, tiles_n;  // 64-pixel tiles over the dz pixel grid
  int log_tw, log_th;
  int num_blocks;                 // total pixel blocks = tiles_w*tiles_h*tiles_n
  int blocks_per_split;
  int cout, cin;                  // logical sizes (cin = padded K per tap)
  int kc_a, ma;                   // dz channel box width (<=64) and number of boxes per 128-row tile
  int kc_b, nb;                   // x channel box width (<=64) and boxes per BN tile
  int bn;                         // cin tile width = kc_b*nb
  int tpc, tap_groups, num_taps;  // taps per CTA, groups, total taps
  int stages;                     // ring depth (kWgStages .. kWgMaxStages)
  int cin_tiles, cout_tiles;
  int dz_c0;                      // channel offset of dz slice inside its buffer
  float* ws;                      // [split][cout][num_taps][cin] fp32
  ConvTap taps[kMaxTaps];         // x taps (c0, dw, p, dh).  ks > 0 (pixel-grouped stem, yb200_conv2d_wgrad_grouped): only the first 16 * ks channels
                                  // of this tap's x box are real (the box starts at the one neighbour pixel that matters, the rest is TMA zero fill):
                                  // the MMA runs with N = 16 * ks and the epilogue stores those columns at input channel kb, zeros elsewhere
};

template <int TMEM_COLS>
__global__ void __launch_bounds__(kConvThreads)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ WgradParams p) {
  pdl_sync();
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2 * kWgMaxStages + 1];
  const int num_stages = p.stages;
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kWgMaxStages]);
  const uint32_t bar_acc = smem_u32(&s_bar[2 * kWgMaxStages]);

  // work decomposition: blockIdx.x = ((cout_tile * cin_tiles + cin_tile) * tap_groups + group), blockIdx.y = split
  int w = blockIdx.x;
  const int group = w % p.tap_groups;
  w /= p.tap_groups;
  const int cin_tile = w % p.cin_tiles;
  const int cout_tile = w / p.cin_tiles;
  const int split = blockIdx.y;
  const int tap0 = group * p.tpc;
  const int ntap = min(p.tpc, p.num_taps - tap0);
  const int blk_begin = split * p.blocks_per_split;
  const int blk_end = min(p.num_blocks, blk_begin + p.blocks_per_split);
  const int nblk = blk_end - blk_begin;

  const uint32_t a_box_bytes = kWgPix * p.kc_a * 2;
  const uint32_t b_box_bytes = kWgPix * p.kc_b * 2;
  const uint32_t a_bytes = a_box_bytes * p.ma;
  const uint32_t b_tap_bytes = b_box_bytes * p.nb;
  const uint32_t stage_bytes = a_bytes + b_tap_bytes * p.tpc;  // constant stride even for a short last group

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmDz);
      tma_prefetch_desc(&tmX);
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        int t = blk_begin + b;
        const int tw = t % p.tiles_w;
        t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int tn = t / p.tiles_h;
        const int w0 = tw << p.log_tw, h0 = th << p.log_th, n0 = tn << (6 - p.log_tw - p.log_th);
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t full = bar_full + 8 * stage;
        mbar_expect_tx(full, a_bytes + b_tap_bytes * ntap);
        for (int j = 0; j < p.ma; ++j)
          tma_load_5d(sa + j * a_box_bytes, &tmDz, full, p.dz_c0 + cout_tile * 128 + j * p.kc_a, w0, 0, h0, n0);
        for (int ti = 0; ti < ntap; ++ti) {
          const ConvTap& tp = p.taps[tap0 + ti];
          const uint32_t sb = sa + a_bytes + ti * b_tap_bytes;
          for (int j = 0; j < p.nb; ++j)
            tma_load_5d(sb + j * b_box_bytes, &tmX, full, tp.c0 + cin_tile * p.bn + j * p.kc_b, w0 + tp.dw, tp.p,
                        h0 + tp.dh, n0);
        }
        if (++stage == num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_full = umma_idesc_bf16(128, p.bn, 1, 1);
      const uint32_t lcode_a = umma_layout_code(p.kc_a * 2);
      const uint32_t lcode_b = umma_layout_code(p.kc_b * 2);
      const uint32_t row_a = p.kc_a * 2, row_b = p.kc_b * 2;
      // MN-major canonical layout: LBO = byte distance between channel boxes, SBO = 8 pixel rows.
      // When the cout tile is narrower than 128 rows the extra rows alias box 0 (LBO 0): they only
      // produce accumulator rows that the epilogue never reads.
      const uint32_t lbo_a = (p.ma * p.kc_a >= 128) ? a_box_bytes : 0u;
      const uint32_t lbo_b = b_box_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        for (int ti = 0; ti < ntap; ++ti) {
          const uint32_t sb = sa + a_bytes + ti * b_tap_bytes;
          const int ks = p.taps[tap0 + ti].ks;
          const uint32_t idesc = ks > 0 ? umma_idesc_bf16(128, 16 * ks, 1, 1) : idesc_full;
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 16 * row_a, lbo_a, 8 * row_a, lcode_a);
            const uint64_t db = umma_smem_desc(sb + k * 16 * row_b, lbo_b, 8 * row_b, lcode_b);
            umma_f16(tmem_base + ti * p.bn, da, db, idesc, (b | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * stage);
        if (++stage == num_stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar_acc);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int co = cout_tile * 128 + row;
    const bool row_ok = (row < p.ma * p.kc_a) && (co < p.cout);
    if (nblk > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    float* wsb = p.ws + (long long)split * p.cout * p.num_taps * p.cin;
    for (int ti = 0; ti < ntap; ++ti) {
      float* dst = wsb + ((long long)co * p.num_taps + tap0 + ti) * p.cin + cin_tile * p.bn;
      const int ks = p.taps[tap0 + ti].ks, shift = p.taps[tap0 + ti].kb;
      for (int c = 0; c < p.bn; c += 16) {
        uint32_t r[16];
        const int src = ks > 0 ? c - shift : c;  // accumulator column of output column c (sparse taps: only [shift, shift + 16 ks) exist)
        if (nblk > 0 && (ks == 0 || (src >= 0 && src < 16 * ks))) {
          tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ti * p.bn + src, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = 0u;
        }
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<uint4*>(dst + c + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// Sum the split-K partials and scatter into the reference's OIHW fp32 gradient layout
// (torch.nn.Conv2d.weight.grad: [Cout][Cin][kh][kw]); padded input channels (ci >= cin_real) are dropped.
// accumulate != 0 adds to the existing gradient (gradient accumulation across micro-batches).
// block = (32 outputs, 8 split lanes): the serial chain over splits is 8x shorter; partials are combined in a fixed order
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ grad_oihw, int splits, int cout, int num_taps, int cin_pad,
                    int cin_real, int accumulate) {
  pdl_sync();
  __shared__ float part[8][33];
  const unsigned total = static_cast<unsigned>(cout) * num_taps * cin_pad;
  const unsigned tc = static_cast<unsigned>(num_taps) * cin_pad;
  const int ox = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (unsigned base = blockIdx.x * 32u; base < total; base += gridDim.x * 32u) {
    const unsigned i = base + ox;
    float acc = 0.f;
    if (i < total) {
      // four independent loads in flight per thread (the kernel is latency bound: one 4-byte load per thread and iteration otherwise);
      // fixed association order, so the result does not depend on the launch
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int s = sl;
      for (; s + 24 < splits; s += 32) {
        a0 += ws[static_cast<size_t>(s) * total + i];
        a1 += ws[static_cast<size_t>(s + 8) * total + i];
        a2 += ws[static_cast<size_t>(s + 16) * total + i];
        a3 += ws[static_cast<size_t>(s + 24) * total + i];
      }
      for (; s < splits; s += 8) a0 += ws[static_cast<size_t>(s) * total + i];
      acc = (a0 + a1) + (a2 + a3);
    }
    part[sl][ox] = acc;
    __syncthreads();
    if (sl == 0 && i < total) {
      const float v = ((part[0][ox] + part[1][ox]) + (part[2][ox] + part[3][ox])) + ((part[4][ox] + part[5][ox]) + (part[6][ox] + part[7][ox]));
      const unsigned co = i / tc;
      const unsigned r = i - co * tc;
      const unsigned tap = r / cin_pad;
      const unsigned ci = r - tap * cin_pad;
      if (ci < static_cast<unsigned>(cin_real)) {
        float* g = grad_oihw + (static_cast<size_t>(co) * cin_real + ci) * num_taps + tap;
        *g = accumulate ? (*g + v) : v;
      }
    }
    __syncthreads();
  }
}

}  // namespace yb

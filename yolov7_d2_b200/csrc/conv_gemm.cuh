// Implicit-GEMM convolution on tcgen05 (sm_100a).
//
//   D[128 pixels, BLOCK_N channels] = sum over (tap, cin-block)  A_tap[128 px, BLOCK_K] * B_tap[BLOCK_N, BLOCK_K]^T
//
// A is the NHWC bf16 activation tensor seen through a 5-D TMA map (c, w, p, h, n): a "tap" is a
// coordinate offset of the 128-pixel box (TW x TH x TN), so zero padding is TMA out-of-bounds fill and
// stride-2 convolutions are taps on the space-to-depth view (p = row parity, column parity folded into c).
// B is the packed weight matrix [rows][taps*K] (K-major).  Accumulators live in TMEM; one CTA computes
// one 128 x BLOCK_N output tile with a TMA->smem->UMMA mbarrier ring.
//
// Used for: conv forward (F3/F2/F8 of SURVEY.md par.8a), data-gradient (same kernel, transposed tap table)
// and the 1x1 prediction convolutions (fp32 + bias epilogue).
#pragma once
#include <cuda_fp16.h>

#include "sm100.cuh"

namespace yb {

constexpr int kConvThreads = 192;  // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-5: epilogue
constexpr int kMaxTaps = 54;  // 9 spatial taps x up to 6 operand-split terms (strict mode, conv_api.cu)
constexpr int kMaxStages = 4;

enum EpiMode : int {
  EPI_BF16 = 0,       // out = bf16(acc [+ addend])                          (data gradients)
  EPI_F16 = 1,        // out = fp16(acc)                                     (pre-BatchNorm conv output, eval mode)
  EPI_F16_STATS = 2,  // out = fp16(acc) + per-channel sum / sum-of-squares of the stored values (fp64 atomics)
  EPI_F32_BIAS = 3,   // out = acc + bias[c]   (fp32, arbitrary element strides)
  EPI_BF16_BN_SILU = 4,  // out = bf16(SiLU(acc*scale[c] + shift[c]) [+ addend]): eval-mode BatchNorm folded into the conv
                         // (the fold of utils/checkpoint.py:11-43 applied as an epilogue), residual added after the activation
  EPI_BF16_AFFINE = 5,   // out = bf16(acc*scale[c] + shift[c] [+ addend]); scale / shift may be null (1 / 0): Linear / Conv bias,
                         // ConvNeXt layer scale + residual (convnext.py:54-59)
  EPI_BF16_BIAS_GELU = 6,  // u = bf16(acc + shift[c]) -> aux_out (optional), out = bf16(GELU(u)), exact erf form (convnext.py:52-53)
  EPI_BF16_GELU_BWD = 7,   // out = bf16(acc * GELU'(u)), u read from aux_in; optional per-channel sums of the stored values -> stat_sum
                           // (gradient of the Linear bias that produced u)
  EPI_BF16_BIAS_RELU = 8,  // out = bf16(max(acc + shift[c], 0)): Linear + ReLU of the transformer FFN (detr_backbone.py:167)
  EPI_BF16_RELU_BWD = 9,   // out = bf16(aux_in > 0 ? acc : 0) (aux_in = the ReLU output); optional column sums -> stat_sum
};

__device__ __forceinline__ bool epi_has_stats(int mode, const double* stat_sum) {
  return mode == EPI_F16_STATS || ((mode == EPI_BF16_GELU_BWD || mode == EPI_BF16_RELU_BWD) && stat_sum != nullptr);
}
// Phi(u) = 0.5 (1 + erf(u / sqrt 2)) through Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the bf16 resolution of the stored
// result): one MUFU.RCP, one MUFU.EX2 and 7 FMAs instead of erff's branchy polynomial -- the GEMMs that carry these epilogues have
// K = C and N = 4C, so they are bound by epilogue instruction issue, not by the tensor pipe.  e = exp(-u^2 / 2) is shared with GELU'.
__device__ __forceinline__ float gelu_phi(float u, float& e) {
  const float z = fabsf(u) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  e = __expf(-z * z);
  const float q = 0.5f * poly * e;
  return u >= 0.f ? 1.f - q : q;
}
__device__ __forceinline__ float gelu_erf(float u) {
  float e;
  return u * gelu_phi(u, e);
}
__device__ __forceinline__ float gelu_erf_grad(float u) {
  float e;
  const float phi = gelu_phi(u, e);
  return fmaf(u * 0.3989422804014327f, e, phi);
}

struct ConvTap {
  int c0;  // coordinate offset in the innermost (channel) dimension of the A map
  int dw;  // offset in w
  int p;   // coordinate in the parity dimension
  int dh;  // offset in h
  int kb;  // column offset of this tap inside the B matrix
  int ks;  // > 0: only the first ks 16-wide K steps of each k-block of this tap carry non-zero operands (persistent kernel: the rest is skipped)
};

// Fused BatchNorm-backward statistics (kernel variant 2, data-gradient launches): the epilogue that produces the FINAL gradient da of an
// activation a = SiLU(BN(z)) also reads z and accumulates  S1 = sum du,  S2 = sum du * z  with du = da * SiLU'(z*scale + shift)  per channel
// -- the reduction pass of BatchNorm backward (dbeta = S1, dgamma = invstd * (S2 - mean * S1)) without a second trip of da and z through HBM.
// Up to two segments: the gradient tensor may be a concat buffer whose channel ranges come from different BatchNorm layers.
struct BnBwdSeg {
  int col_begin, col_end;      // columns (channels of the gradient tensor being written) covered by this segment; multiples of 32
  const __half* z;             // the producer's pre-BN output, pointing at (pixel 0, channel col_begin)
  long long z_sn, z_sh, z_sw;  // element strides of z over (image, row, column); z has the geometry of the gradient tensor
  const float* scale;          // BatchNorm scale / shift of the producer, indexed by (column - col_begin)
  const float* shift;
  double* sum_du;              // S1 accumulator, indexed by (column - col_begin)
  double* sum_duz;             // S2 accumulator
};

struct ConvGemmParams {
  int num_bnseg;
  BnBwdSeg bnseg[2];
  int tiles_w, tiles_h, tiles_n;
  int log_tw, log_th;          // tile = (1<<log_tw) x (1<<log_th) x (128 >> (log_tw+log_th)) pixels
  int num_taps, cin_blocks;    // K loop = num_taps * cin_blocks blocks of BLOCK_K
  int n_valid, h_valid, w_valid;  // pixel-grid extents (tile-space); pixels outside are neither stored nor counted
  int cout;                    // valid output channels (columns >= cout are dropped)
  int epi_mode;
  // output element (n, y, x, c) lives at out[n*out_sn + (y*out_mh+out_ph)*out_sh + (x*out_mw+out_pw)*out_sw + c*out_sc]
  long long out_sn, out_sh, out_sw;
  int out_sc, out_mh, out_ph, out_mw, out_pw;
  void* out;
  const __nv_bfloat16* addend;  // optional, bf16, same (n,y,x) -> offset mapping with its own strides, channel stride 1
  long long add_sn, add_sh, add_sw;
  const float* bias;
  const float* scale;  // EPI_BF16_BN_SILU: per-channel scale / shift (gamma/sqrt(var+eps), beta - mean*scale)
  const float* shift;
  double* stat_sum;
  double* stat_sq;
  int stat_fold;  // > 0: column c accumulates into statistic c % stat_fold (pixel-grouped views: several columns are the same channel)
  // num_phases == 4: ONE launch computes the four output-parity classes of a stride-2 data gradient.  Work item wi = 4 * tile + phase;
  // phase (ph, pw) = (wi >> 1 & 1, wi & 1) uses taps[phase_tap[phase] .. phase_tap[phase + 1]) and writes pixel (2y + ph, 2x + pw).  The four
  // phases of a tile run on neighbouring CTAs at about the same time, so the gradient tile they share comes from HBM once (four separate
  // launches read the whole tensor four times).
  int num_phases;
  int phase_tap[5];
  int b_img_rows; // > 0: per-image weights -- the tile of image n reads weight rows n * b_img_rows + column (tiles must not span images)
  int xpose;      // EPI_F32_BIAS with channel-contiguous rows: transpose each 32 x 32 chunk through shared memory (kXposeBytes behind the ring)
  const __nv_bfloat16* aux_in;  // EPI_BF16_GELU_BWD: pre-activation u, same geometry as `out`
  __nv_bfloat16* aux_out;       // EPI_BF16_BIAS_GELU: where u is stored (may be null), same geometry as `out`
  ConvTap taps[kMaxTaps];
};

__device__ __forceinline__ void conv_decode_work(const ConvGemmParams& p, int wi, int& m, int& tap0, int& ntaps, int& oph, int& opw) {
  if (p.num_phases == 4) {
    m = wi >> 2;
    const int ph = wi & 3;
    tap0 = p.phase_tap[ph];
    ntaps = p.phase_tap[ph + 1] - tap0;
    oph = ph >> 1;
    opw = ph & 1;
  } else {
    m = wi; tap0 = 0; ntaps = p.num_taps; oph = p.out_ph; opw = p.out_pw;
  }
}

template <int BLOCK_N, int BLOCK_K>
struct ConvGemmCfg {
  static constexpr int kSwizzle = BLOCK_K * 2;  // bytes of one K-row: 32 / 64 / 128
  static constexpr int kABytes = 128 * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int kChunk = BLOCK_N < 32 ? 16 : 32;  // epilogue column chunk
  static_assert(BLOCK_K == 16 || BLOCK_K == 32 || BLOCK_K == 64, "BLOCK_K");
  static_assert(BLOCK_N == 16 || BLOCK_N == 32 || BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  static_assert(kABytes % (8 * kSwizzle) == 0 && kStageBytes % (8 * kSwizzle) == 0, "tiles must start on a swizzle-pattern boundary");
};

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store_f16x8(__half* dst, const float* v) {
  uint4 u;
  u.x = pack_f16x2(v[0], v[1]);
  u.y = pack_f16x2(v[2], v[3]);
  u.z = pack_f16x2(v[4], v[5]);
  u.w = pack_f16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}

// One chunk (CH columns of one accumulator row per lane) of the epilogue, shared by all three kernels.
//   part_sum / part_sq: this warp's shared-memory slots for the chunk's columns (EPI_F16_STATS); `accumulate` adds to them
//   (persistent kernels: statistics over all tiles of the CTA) instead of overwriting.
// Per-column parameters (scale | shift / bias) of the CTA's column tile are staged in shared memory once per CTA; the epilogue reads
// them with 16-byte broadcast loads (per-element __ldg plus a null test per element made the parameterised epilogues 2-3x slower
// than the plain one).  Columns beyond cout hold (1, 0).
template <int BLOCK_N>
__device__ __forceinline__ void stage_col_params(const ConvGemmParams& p, int col0, float (*s_col)[BLOCK_N]) {
  const float* sh = p.shift ? p.shift : p.bias;
  for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) {
    const bool ok = col0 + i < p.cout;
    s_col[0][i] = (ok && p.scale) ? p.scale[col0 + i] : 1.f;
    s_col[1][i] = (ok && sh) ? sh[col0 + i] : 0.f;
  }
}

__device__ __forceinline__ int bnseg_of(const ConvGemmParams& p, int col) {
  if (p.num_bnseg > 0 && col >= p.bnseg[0].col_begin && col < p.bnseg[0].col_end) return 0;
  if (p.num_bnseg > 1 && col >= p.bnseg[1].col_begin && col < p.bnseg[1].col_end) return 1;
  return -1;
}
template <int BLOCK_N>
__device__ __forceinline__ void stage_col_params_bnseg(const ConvGemmParams& p, int col0, float (*s_col)[BLOCK_N]) {
  for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) {
    const int s = bnseg_of(p, col0 + i);
    s_col[0][i] = s >= 0 ? p.bnseg[s].scale[col0 + i - p.bnseg[s].col_begin] : 1.f;
    s_col[1][i] = s >= 0 ? p.bnseg[s].shift[col0 + i - p.bnseg[s].col_begin] : 0.f;
  }
}
__device__ __forceinline__ float silu_grad_f(float u, float d) {  // d * d/du [u * sigmoid(u)], sigmoid on MUFU.TANH as elementwise.cu
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
  const float sg = fmaf(t, 0.5f, 0.5f);
  return (d * sg) * fmaf(u, 1.f - sg, 1.f);
}
// fp16 z chunk of one pixel row (CH channels) for the fused BatchNorm-backward statistics
template <int CH>
__device__ __forceinline__ void load_z_chunk(const __half* zrow, bool valid, uint4 (&dst)[CH / 8]) {
#pragma unroll
  for (int k = 0; k < CH / 8; ++k) {
    dst[k] = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
      const void* ptr = zrow + 8 * k;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(dst[k].x), "=r"(dst[k].y), "=r"(dst[k].z), "=r"(dst[k].w) : "l"(ptr));
    }
  }
}

// The epilogue's per-element side input (GELU_BWD: the pre-activation u; otherwise the addend / residual) is a dependent global load in
// a warp that has nothing else to do: the persistent kernels fetch it one chunk ahead (before the accumulator barrier for the first
// chunk of a tile) so that its latency overlaps the TMEM load and the arithmetic of the previous chunk.
template <int CH>
__device__ __forceinline__ void load_side_chunk(const __nv_bfloat16* side_row, bool valid, int cbase, int cout, uint4 (&dst)[CH / 8]) {
#pragma unroll
  for (int k = 0; k < CH / 8; ++k) {
    dst[k] = make_uint4(0u, 0u, 0u, 0u);
    if (valid && cbase + 8 * k < cout) dst[k] = *reinterpret_cast<const uint4*>(side_row + cbase + 8 * k);
  }
}

// EXT = false: the YOLOX training / inference modes only (EPI_BF16 .. EPI_BF16_BN_SILU); EXT = true adds the ConvNeXt / transformer
// modes and the prefetched side input.  Two instantiations per kernel keep the hot YOLOX kernels as small as they were before the
// extra modes existed (the combined epilogue cost the 3x3 kernels 2-13 % through registers and code size).
// fp32 rows with a pitch that is not a multiple of 16 bytes (the [B, A, 85] prediction tensor): a lane-per-pixel store touches 32 different
// sectors per instruction.  With the per-warp scratch each store instruction writes 32 consecutive floats of ONE pixel row instead.
constexpr int kXposeWarpFloats = 32 * 33 + 64;              // 32 x 32 chunk (pitch 33: conflict free both ways) + 32 row offsets (8 B each)
constexpr int kXposeBytes = 8 * kXposeWarpFloats * 4;       // eight epilogue warps

template <int CH, bool EXT = true>
__device__ __forceinline__ void conv_epilogue_chunk(const ConvGemmParams& p, float (&v)[CH], bool valid, long long pix_off, long long add_off,
                                                    int cbase, int lane, float* part_sum, float* part_sq, bool accumulate,
                                                    const float* col_scale, const float* col_shift, const uint4* side = nullptr,
                                                    const uint4* zq = nullptr, float* xp = nullptr) {
  if (p.epi_mode == EPI_F32_BIAS) {
    if (xp != nullptr) {
      long long* s_off = reinterpret_cast<long long*>(xp + 32 * 33);
#pragma unroll
      for (int i = 0; i < CH; ++i) xp[lane * 33 + i] = v[i] + col_shift[i];
      s_off[lane] = valid ? pix_off : -1;
      __syncwarp();
      float* o = reinterpret_cast<float*>(p.out) + cbase + lane;
      const bool colok = lane < CH && cbase + lane < p.cout;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const long long off = s_off[r];
        if (off >= 0 && colok) o[off] = xp[r * 33 + lane];
      }
      __syncwarp();
      return;
    }
    if (valid) {
      float* o = reinterpret_cast<float*>(p.out) + pix_off;
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (cbase + i < p.cout) o[(long long)(cbase + i) * p.out_sc] = v[i] + col_shift[i];
    }
    return;
  }
  if (p.epi_mode == EPI_BF16_BN_SILU) {
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
      const float4 sc = *reinterpret_cast<const float4*>(col_scale + i), sh = *reinterpret_cast<const float4*>(col_shift + i);
      const float u0 = fmaf(v[i], sc.x, sh.x), u1 = fmaf(v[i + 1], sc.y, sh.y), u2 = fmaf(v[i + 2], sc.z, sh.z), u3 = fmaf(v[i + 3], sc.w, sh.w);
      // rounded before the residual add, as a materialised activation
      v[i] = bf16_round(__fdividef(u0, 1.f + __expf(-u0)));
      v[i + 1] = bf16_round(__fdividef(u1, 1.f + __expf(-u1)));
      v[i + 2] = bf16_round(__fdividef(u2, 1.f + __expf(-u2)));
      v[i + 3] = bf16_round(__fdividef(u3, 1.f + __expf(-u3)));
    }
  }
  if constexpr (EXT) {
    if (p.epi_mode == EPI_BF16_AFFINE) {
      if (p.scale != nullptr) {
  #pragma unroll
        for (int i = 0; i < CH; i += 4) {
          const float4 sc = *reinterpret_cast<const float4*>(col_scale + i);
          v[i] *= sc.x; v[i + 1] *= sc.y; v[i + 2] *= sc.z; v[i + 3] *= sc.w;
        }
      }
      if (p.shift != nullptr) {
  #pragma unroll
        for (int i = 0; i < CH; i += 4) {
          const float4 sh = *reinterpret_cast<const float4*>(col_shift + i);
          v[i] += sh.x; v[i + 1] += sh.y; v[i + 2] += sh.z; v[i + 3] += sh.w;
        }
      }
    }
    if (p.epi_mode == EPI_BF16_BIAS_RELU) {
  #pragma unroll
      for (int i = 0; i < CH; i += 4) {
        const float4 sh = *reinterpret_cast<const float4*>(col_shift + i);
        v[i] = fmaxf(v[i] + sh.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + sh.y, 0.f);
        v[i + 2] = fmaxf(v[i + 2] + sh.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + sh.w, 0.f);
      }
    }
    if (p.epi_mode == EPI_BF16_RELU_BWD && valid) {
      const __nv_bfloat16* a = p.aux_in + pix_off + cbase;
  #pragma unroll
      for (int i = 0; i < CH; i += 8) {
        if (cbase + i < p.cout) {
          const uint4 u = side ? side[i >> 3] : *reinterpret_cast<const uint4*>(a + i);
          // bf16 sign / zero test on the raw bits: positive and non-zero
          v[i + 0] = (u.x & 0xFFFFu) - 1u < 0x7FFFu ? v[i + 0] : 0.f; v[i + 1] = (u.x >> 16) - 1u < 0x7FFFu ? v[i + 1] : 0.f;
          v[i + 2] = (u.y & 0xFFFFu) - 1u < 0x7FFFu ? v[i + 2] : 0.f; v[i + 3] = (u.y >> 16) - 1u < 0x7FFFu ? v[i + 3] : 0.f;
          v[i + 4] = (u.z & 0xFFFFu) - 1u < 0x7FFFu ? v[i + 4] : 0.f; v[i + 5] = (u.z >> 16) - 1u < 0x7FFFu ? v[i + 5] : 0.f;
          v[i + 6] = (u.w & 0xFFFFu) - 1u < 0x7FFFu ? v[i + 6] : 0.f; v[i + 7] = (u.w >> 16) - 1u < 0x7FFFu ? v[i + 7] : 0.f;
        }
      }
    }
    if (p.epi_mode == EPI_BF16_BIAS_GELU) {
  #pragma unroll
      for (int i = 0; i < CH; i += 4) {
        const float4 sh = *reinterpret_cast<const float4*>(col_shift + i);
        v[i] = bf16_round(v[i] + sh.x); v[i + 1] = bf16_round(v[i + 1] + sh.y);
        v[i + 2] = bf16_round(v[i + 2] + sh.z); v[i + 3] = bf16_round(v[i + 3] + sh.w);
      }
      if (p.aux_out != nullptr && valid) {
        __nv_bfloat16* o = p.aux_out + pix_off + cbase;
  #pragma unroll
        for (int i = 0; i < CH; i += 8)
          if (cbase + i < p.cout) store_bf16x8(o + i, v + i);
      }
  #pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = gelu_erf(v[i]);
    }
    if (p.epi_mode == EPI_BF16_GELU_BWD && valid) {
      const __nv_bfloat16* a = p.aux_in + pix_off + cbase;
  #pragma unroll
      for (int i = 0; i < CH; i += 8) {
        if (cbase + i < p.cout) {
          const uint4 u = side ? side[i >> 3] : *reinterpret_cast<const uint4*>(a + i);
          v[i + 0] *= gelu_erf_grad(bf16_lo(u.x)); v[i + 1] *= gelu_erf_grad(bf16_hi(u.x));
          v[i + 2] *= gelu_erf_grad(bf16_lo(u.y)); v[i + 3] *= gelu_erf_grad(bf16_hi(u.y));
          v[i + 4] *= gelu_erf_grad(bf16_lo(u.z)); v[i + 5] *= gelu_erf_grad(bf16_hi(u.z));
          v[i + 6] *= gelu_erf_grad(bf16_lo(u.w)); v[i + 7] *= gelu_erf_grad(bf16_hi(u.w));
        }
      }
    }
  }
  if (p.addend != nullptr && valid) {
    const __nv_bfloat16* a = p.addend + add_off + cbase;
#pragma unroll
    for (int i = 0; i < CH; i += 8) {
      if (cbase + i < p.cout) {
        const uint4 u = (EXT && side && p.epi_mode != EPI_BF16_GELU_BWD && p.epi_mode != EPI_BF16_RELU_BWD) ? side[i >> 3] : *reinterpret_cast<const uint4*>(a + i);
        v[i + 0] += bf16_lo(u.x); v[i + 1] += bf16_hi(u.x);
        v[i + 2] += bf16_lo(u.y); v[i + 3] += bf16_hi(u.y);
        v[i + 4] += bf16_lo(u.z); v[i + 5] += bf16_hi(u.z);
        v[i + 6] += bf16_lo(u.w); v[i + 7] += bf16_hi(u.w);
      }
    }
  }
  // round once (one packed conversion per two values; the kernels that carry this epilogue are bound by its instruction count on the narrow
  // layers, ncu: ~50 % issue-slot utilisation with 8 epilogue warps per CTA); statistics describe exactly the values that are stored
  const bool f16 = p.epi_mode == EPI_F16 || p.epi_mode == EPI_F16_STATS;
  const bool need_vals = p.epi_mode == EPI_F16_STATS || zq != nullptr || (EXT && p.stat_sum != nullptr);
  if (f16) {
    uint32_t pk[CH / 2];
#pragma unroll
    for (int i = 0; i < CH; i += 2) pk[i >> 1] = pack_f16x2(v[i], v[i + 1]);
    if (valid) {
      __half* o = reinterpret_cast<__half*>(p.out) + pix_off + cbase;
#pragma unroll
      for (int i = 0; i < CH; i += 8)
        if (cbase + i < p.cout) *reinterpret_cast<uint4*>(o + i) = make_uint4(pk[i >> 1], pk[(i >> 1) + 1], pk[(i >> 1) + 2], pk[(i >> 1) + 3]);
    }
    if (need_vals) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) {
        __half2 h;
        *reinterpret_cast<uint32_t*>(&h) = pk[i >> 1];
        const float2 f = __half22float2(h);
        v[i] = valid ? f.x : 0.f;
        v[i + 1] = valid ? f.y : 0.f;
      }
    }
  } else {
    uint32_t pk[CH / 2];
#pragma unroll
    for (int i = 0; i < CH; i += 2) pk[i >> 1] = pack_bf16x2(v[i], v[i + 1]);
    if (valid) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix_off + cbase;
#pragma unroll
      for (int i = 0; i < CH; i += 8)
        if (cbase + i < p.cout) *reinterpret_cast<uint4*>(o + i) = make_uint4(pk[i >> 1], pk[(i >> 1) + 1], pk[(i >> 1) + 2], pk[(i >> 1) + 3]);
    }
    if (need_vals) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) {
        v[i] = valid ? bf16_lo(pk[i >> 1]) : 0.f;
        v[i + 1] = valid ? bf16_hi(pk[i >> 1]) : 0.f;
      }
    }
  }
  if (zq != nullptr) {
    // fused BatchNorm-backward statistics on the values just stored (v = rounded da; 0 on masked rows): du -> v, du * z -> zf
    float zf[CH];
#pragma unroll
    for (int k = 0; k < CH / 8; ++k) {
      const uint32_t w4[4] = {zq[k].x, zq[k].y, zq[k].z, zq[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 h;
        *reinterpret_cast<uint32_t*>(&h) = w4[j];
        const float2 f = __half22float2(h);
        zf[8 * k + 2 * j] = f.x;
        zf[8 * k + 2 * j + 1] = f.y;
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float du = silu_grad_f(fmaf(zf[i], col_scale[i], col_shift[i]), v[i]);
      v[i] = du;
      zf[i] *= du;
    }
    float cs, cq;
    if constexpr (CH == 32) { cs = warp_colsum32(v, lane); cq = warp_colsum32(zf, lane); }
    else { cs = warp_colsum16(v, lane); cq = warp_colsum16(zf, lane); }
    if (lane < CH) {
      part_sum[lane] = accumulate ? part_sum[lane] + cs : cs;
      part_sq[lane] = accumulate ? part_sq[lane] + cq : cq;
    }
    return;
  }
  if (p.epi_mode == EPI_F16_STATS) {
    float sq[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) sq[i] = v[i] * v[i];
    float cs, cq;
    if constexpr (CH == 32) { cs = warp_colsum32(v, lane); cq = warp_colsum32(sq, lane); }
    else { cs = warp_colsum16(v, lane); cq = warp_colsum16(sq, lane); }
    if (lane < CH) {  // each (warp, column) slot has exactly one writer
      part_sum[lane] = accumulate ? part_sum[lane] + cs : cs;
      part_sq[lane] = accumulate ? part_sq[lane] + cq : cq;
    }
  } else if (EXT && (p.epi_mode == EPI_BF16_GELU_BWD || p.epi_mode == EPI_BF16_RELU_BWD) && p.stat_sum != nullptr) {
    float cs;
    if constexpr (CH == 32) cs = warp_colsum32(v, lane);
    else cs = warp_colsum16(v, lane);
    if (lane < CH) part_sum[lane] = accumulate ? part_sum[lane] + cs : cs;
  }
}

template <int BLOCK_N, int BLOCK_K>
__global__ void __launch_bounds__(kConvThreads)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ ConvGemmParams p, int num_stages) {
  pdl_sync();
  using Cfg = ConvGemmCfg<BLOCK_N, BLOCK_K>;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2 * kMaxStages + 1];
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[4][2][BLOCK_N];  // per epilogue warp: column sums / sums of squares of its 32 rows
  __shared__ __align__(16) float s_col[2][BLOCK_N];  // per-column scale | shift of this column tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // swizzled tiles need 1024B alignment
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kMaxStages]);
  const uint32_t bar_acc = smem_u32(&s_bar[2 * kMaxStages]);

  // tile coordinates
  int t = blockIdx.x;
  const int tw = t % p.tiles_w;
  t /= p.tiles_w;
  const int th = t % p.tiles_h;
  const int tn = t / p.tiles_h;
  const int log_tw = p.log_tw, log_th = p.log_th;
  const int w0 = tw << log_tw, h0 = th << log_th, n0 = tn << (7 - log_tw - log_th);
  const int col0 = blockIdx.y * BLOCK_N;
  const int num_kb = p.num_taps * p.cin_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  stage_col_params<BLOCK_N>(p, col0, s_col);
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(smem_u32(&s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      int tap = 0, cb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kABytes;
        const uint32_t full = bar_full + 8 * stage;
        mbar_expect_tx(full, Cfg::kStageBytes);
        const ConvTap& tp = p.taps[tap];
        tma_load_5d(sa, &tmA, full, tp.c0 + cb * BLOCK_K, w0 + tp.dw, tp.p, h0 + tp.dh, n0);
        tma_load_2d(sb, &tmB, full, tp.kb + cb * BLOCK_K, col0);
        if (++cb == p.cin_blocks) { cb = 0; ++tap; }
        if (++stage == num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, 0);
      constexpr uint32_t lcode = umma_layout_code(Cfg::kSwizzle);
      constexpr uint32_t sbo = 8 * Cfg::kSwizzle;  // 8 rows of one swizzle atom
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t da = umma_smem_desc(sa + k * 32, 16, sbo, lcode);
          const uint64_t db = umma_smem_desc(sb + k * 32, 16, sbo, lcode);
          umma_f16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
        if (++stage == num_stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar_acc);  // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;
    const int xl = m & ((1 << log_tw) - 1);
    const int yl = (m >> log_tw) & ((1 << log_th) - 1);
    const int nl = m >> (log_tw + log_th);
    const int x = w0 + xl, y = h0 + yl, n = n0 + nl;
    const bool valid = (x < p.w_valid) && (y < p.h_valid) && (n < p.n_valid);
    const long long pix_off = (long long)n * p.out_sn + (long long)(y * p.out_mh + p.out_ph) * p.out_sh +
                              (long long)(x * p.out_mw + p.out_pw) * p.out_sw;
    const long long add_off = (long long)n * p.add_sn + (long long)(y * p.out_mh + p.out_ph) * p.add_sh +
                              (long long)(x * p.out_mw + p.out_pw) * p.add_sw;

    mbar_wait(bar_acc, 0);
    tc_fence_after();

    constexpr int CH = Cfg::kChunk;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += CH) {
      uint32_t r[CH];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c;
      if constexpr (CH == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
      tmem_ld_wait();
      float v[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]);
      const int cbase = col0 + c;
      conv_epilogue_chunk<CH>(p, v, valid, pix_off, add_off, cbase, lane, &s_part[q][0][c], &s_part[q][1][c], false, &s_col[0][c], &s_col[1][c]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  if (epi_has_stats(p.epi_mode, p.stat_sum)) {
    // one fp64 atomic per channel and CTA (after the block-wide barrier above: no named barrier, no shared atomics)
    for (int e = threadIdx.x; e < BLOCK_N && col0 + e < p.cout; e += blockDim.x) {  // BLOCK_N may exceed the 192 threads
      const float s1 = (s_part[0][0][e] + s_part[1][0][e]) + (s_part[2][0][e] + s_part[3][0][e]);
      const float s2 = (s_part[0][1][e] + s_part[1][1][e]) + (s_part[2][1][e] + s_part[3][1][e]);
      atomicAdd(p.stat_sum + col0 + e, static_cast<double>(s1));
      if (p.stat_sq != nullptr) atomicAdd(p.stat_sq + col0 + e, static_cast<double>(s2));
    }
  }
}

// ================================================================================================
// Persistent variant: each CTA walks a strided list of 128-pixel tiles for ONE column tile.  The TMA ring runs ahead across
// tile boundaries, accumulators are double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1, per-channel statistics are accumulated in shared memory over all tiles of the CTA (one fp64 atomic per
// channel and CTA instead of per tile), and TMEM allocation / barrier setup / descriptor prefetch are paid once.
// grid.x = n_tiles * groups;  CTA b: column tile b % n_tiles, pixel tiles (b / n_tiles) + i * groups.
// ================================================================================================
constexpr int kMaxStagesP = 8;     // ring slots; one slot holds kb_per_slot consecutive k-blocks under a single mbarrier
constexpr int kConvThreadsP = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue: two warps per TMEM lane quadrant, each
                                    // draining half of the accumulator columns (memory-bound layers are epilogue-bound)

// VAR: 0 = YOLOX training / inference epilogues, 1 = + ConvNeXt / transformer epilogues (EXT), 2 = data gradient with fused
// BatchNorm-backward statistics (BnBwdSeg; epi_mode EPI_BF16)
template <int BLOCK_N, int BLOCK_K, int VAR>
__global__ void __launch_bounds__(kConvThreadsP, BLOCK_N == 256 ? 1 : 2)
conv_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ ConvGemmParams p, int num_stages, int kb_per_slot, int n_tiles, int m_tiles) {
  using Cfg = ConvGemmCfg<BLOCK_N, BLOCK_K>;
  constexpr bool EXT = VAR == 1;
  constexpr bool BNB = VAR == 2;
  constexpr int kAccCols = Cfg::kTmemCols;          // columns of one accumulator
  constexpr int kTmemAlloc = 2 * kAccCols;          // double buffered (power of two >= 64)
  constexpr int kEpiHalves = BLOCK_N >= 32 ? 2 : 1; // column halves, one epilogue warp group (4 warps) each
  constexpr int kHalfCols = BLOCK_N / kEpiHalves;
  constexpr int CH = kHalfCols < 32 ? 16 : 32;      // columns per tcgen05.ld
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2 * kMaxStagesP + 4];
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[4][2][BLOCK_N];
  __shared__ __align__(16) float s_col[2][BLOCK_N];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kMaxStagesP]);
  const uint32_t bar_acc_full = smem_u32(&s_bar[2 * kMaxStagesP]);       // [2]
  const uint32_t bar_acc_empty = smem_u32(&s_bar[2 * kMaxStagesP + 2]);  // [2]

  const int n_tile = blockIdx.x % n_tiles;
  const int group = blockIdx.x / n_tiles;
  const int groups = gridDim.x / n_tiles;
  const int col0 = n_tile * BLOCK_N;
  const int log_tw = p.log_tw, log_th = p.log_th;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, 4 * kEpiHalves);  // one arrival per active epilogue warp
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 4 * 2 * BLOCK_N; i += blockDim.x) (&s_part[0][0][0])[i] = 0.f;
  if (warp == 1) tmem_alloc<kTmemAlloc>(smem_u32(&s_tmem));
  pdl_sync();  // everything above touches only this CTA's shared / tensor memory: it overlaps the tail of the previous kernel
  if constexpr (BNB) stage_col_params_bnseg<BLOCK_N>(p, col0, s_col);
  else stage_col_params<BLOCK_N>(p, col0, s_col);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = group; wi < m_tiles; wi += groups) {  // m_tiles counts work items (pixel tiles x phases)
        int m, tap, ntaps, oph, opw;
        conv_decode_work(p, wi, m, tap, ntaps, oph, opw);
        const int num_kb = ntaps * p.cin_blocks;
        int t = m;
        const int tw = t % p.tiles_w;
        t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int tn = t / p.tiles_h;
        const int w0 = tw << log_tw, h0 = th << log_th, n0 = tn << (7 - log_tw - log_th);
        int cb = 0;
        for (int kb = 0; kb < num_kb; kb += kb_per_slot) {  // num_kb is a multiple of kb_per_slot
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, Cfg::kStageBytes * kb_per_slot);
          for (int j = 0; j < kb_per_slot; ++j) {
            const uint32_t sa = smem_base + (stage * kb_per_slot + j) * Cfg::kStageBytes;
            const ConvTap& tp = p.taps[tap];
            tma_load_5d(sa, &tmA, full, tp.c0 + cb * BLOCK_K, w0 + tp.dw, tp.p, h0 + tp.dh, n0);
            tma_load_2d(sa + Cfg::kABytes, &tmB, full, tp.kb + cb * BLOCK_K, col0 + n0 * p.b_img_rows);
            if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          }
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, 0);
      constexpr uint32_t lcode = umma_layout_code(Cfg::kSwizzle);
      constexpr uint32_t sbo = 8 * Cfg::kSwizzle;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int wi = group; wi < m_tiles; wi += groups, ++it) {
        int m, tap0, ntaps, oph, opw;
        conv_decode_work(p, wi, m, tap0, ntaps, oph, opw);
        const int num_kb = ntaps * p.cin_blocks;
        const int acc = it & 1;
        mbar_wait(bar_acc_empty + 8 * acc, ((it >> 1) & 1) ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * kAccCols;
        int tap = tap0, cb = 0;
        for (int kb = 0; kb < num_kb; kb += kb_per_slot) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          for (int j = 0; j < kb_per_slot; ++j) {
            const uint32_t sa = smem_base + (stage * kb_per_slot + j) * Cfg::kStageBytes;
            const uint32_t sb = sa + Cfg::kABytes;
            const int ks = p.taps[tap].ks > 0 ? p.taps[tap].ks : BLOCK_K / 16;
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              if (k < ks) {
                const uint64_t da = umma_smem_desc(sa + k * 32, 16, sbo, lcode);
                const uint64_t db = umma_smem_desc(sb + k * 32, 16, sbo, lcode);
                umma_f16(tacc, da, db, idesc, (kb | j | k) != 0 ? 1u : 0u);
              }
            }
            if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_acc_full + 8 * acc);
      }
    }
  } else if ((warp - 2) / 4 < kEpiHalves) {
    // ===================== epilogue =====================
    const int q = warp & 3;              // TMEM lane quadrant (must be warp_id % 4)
    const int half = (warp - 2) >> 2;    // which half of the columns this warp drains
    const int cbeg = half * kHalfCols, cend = cbeg + kHalfCols;
    // chunks that start at or beyond the last valid output channel are dead (cout not a multiple of the column tile): skip them
    const int cend_live = min(cend, ((p.cout - col0 + CH - 1) / CH) * CH);
    const bool side_is_aux = p.epi_mode == EPI_BF16_GELU_BWD || p.epi_mode == EPI_BF16_RELU_BWD;
    const __nv_bfloat16* side_base = (EXT && BLOCK_N == 256) ? (side_is_aux ? p.aux_in : p.addend) : nullptr;  // narrower tiles run 2 CTAs / SM at 96 registers: no room
    const int mrow = q * 32 + lane;
    const int xl = mrow & ((1 << log_tw) - 1);
    const int yl = (mrow >> log_tw) & ((1 << log_th) - 1);
    const int nl = mrow >> (log_tw + log_th);
    float* xp = nullptr;
    if (!BNB && p.xpose)
      xp = reinterpret_cast<float*>(smem_dyn + (smem_base - smem_u32(smem_dyn)) + num_stages * kb_per_slot * Cfg::kStageBytes) + (warp - 2) * kXposeWarpFloats;
    int it = 0;
    for (int wi = group; wi < m_tiles; wi += groups, ++it) {
      int m, tap0, ntaps, oph, opw;
      conv_decode_work(p, wi, m, tap0, ntaps, oph, opw);
      int t = m;
      const int tw = t % p.tiles_w;
      t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int tn = t / p.tiles_h;
      const int x = (tw << log_tw) + xl, y = (th << log_th) + yl, n = (tn << (7 - log_tw - log_th)) + nl;
      const bool valid = (x < p.w_valid) && (y < p.h_valid) && (n < p.n_valid);
      const long long pix_off = (long long)n * p.out_sn + (long long)(y * p.out_mh + oph) * p.out_sh +
                                (long long)(x * p.out_mw + opw) * p.out_sw;
      const long long add_off = (long long)n * p.add_sn + (long long)(y * p.out_mh + oph) * p.add_sh +
                                (long long)(x * p.out_mw + opw) * p.add_sw;
      const int acc = it & 1;
      const __nv_bfloat16* side_row = side_base ? side_base + (side_is_aux ? pix_off : add_off) : nullptr;
      uint4 side[CH / 8];
      if constexpr (EXT && BLOCK_N == 256) {
        if (side_base != nullptr && cend_live > cbeg) load_side_chunk<CH>(side_row, valid, col0 + cbeg, p.cout, side);  // in flight during the barrier wait
      }
      // fused BatchNorm-backward statistics: this pixel's row of z in each segment
      const __half* zrow[2] = {nullptr, nullptr};
      if constexpr (BNB) {
#pragma unroll
        for (int sg = 0; sg < 2; ++sg)
          if (sg < p.num_bnseg)
            zrow[sg] = p.bnseg[sg].z + (long long)n * p.bnseg[sg].z_sn + (long long)(y * p.out_mh + oph) * p.bnseg[sg].z_sh +
                       (long long)(x * p.out_mw + opw) * p.bnseg[sg].z_sw - p.bnseg[sg].col_begin;
      }
      mbar_wait(bar_acc_full + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      if (cend_live <= cbeg) {  // every column of this warp's share lies beyond cout: nothing to read
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        continue;
      }
#pragma unroll 1
      for (int c = cbeg; c < cend_live; c += CH) {
        uint32_t r[CH];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + c;
        if constexpr (CH == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
        uint4 side_next[CH / 8];
        const bool more = side_base != nullptr && c + CH < cend_live;
        if constexpr (EXT && BLOCK_N == 256) {  // registers to spare (one CTA per SM): fetch the next chunk's side input before using this one
          if (more) load_side_chunk<CH>(side_row, valid, col0 + c + CH, p.cout, side_next);
        }
        uint4 zq[CH / 8];
        int zseg = -1;
        if constexpr (BNB) {
          zseg = bnseg_of(p, col0 + c);  // warp-uniform: chunks never straddle a segment boundary
          if (zseg >= 0) load_z_chunk<CH>(zrow[zseg] + col0 + c, valid, zq);  // in flight while the accumulator chunk arrives
        }
        tmem_ld_wait();
        if (c + CH >= cend_live) {  // this warp's last chunk is in registers: hand its share of the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        }
        float v[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]);
        const int cbase = col0 + c;
        conv_epilogue_chunk<CH, EXT>(p, v, valid, pix_off, add_off, cbase, lane, &s_part[q][0][c], &s_part[q][1][c], true, &s_col[0][c], &s_col[1][c],
                                (EXT && BLOCK_N == 256 && side_base) ? side : nullptr, (BNB && zseg >= 0) ? zq : nullptr, xp);
        if constexpr (EXT && BLOCK_N == 256) {
          if (more) {
#pragma unroll
            for (int k = 0; k < CH / 8; ++k) side[k] = side_next[k];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemAlloc>(tmem_base);
  if constexpr (BNB) {
    for (int e = threadIdx.x; e < BLOCK_N && col0 + e < p.cout; e += blockDim.x) {
      const int sg = bnseg_of(p, col0 + e);
      if (sg < 0) continue;
      const float s1 = (s_part[0][0][e] + s_part[1][0][e]) + (s_part[2][0][e] + s_part[3][0][e]);
      const float s2 = (s_part[0][1][e] + s_part[1][1][e]) + (s_part[2][1][e] + s_part[3][1][e]);
      atomicAdd(p.bnseg[sg].sum_du + col0 + e - p.bnseg[sg].col_begin, static_cast<double>(s1));
      atomicAdd(p.bnseg[sg].sum_duz + col0 + e - p.bnseg[sg].col_begin, static_cast<double>(s2));
    }
    return;
  }
  if (EXT ? epi_has_stats(p.epi_mode, p.stat_sum) : p.epi_mode == EPI_F16_STATS) {
    for (int e = threadIdx.x; e < BLOCK_N && col0 + e < p.cout; e += blockDim.x) {  // BLOCK_N may exceed the 192 threads
      const float s1 = (s_part[0][0][e] + s_part[1][0][e]) + (s_part[2][0][e] + s_part[3][0][e]);
      const float s2 = (s_part[0][1][e] + s_part[1][1][e]) + (s_part[2][1][e] + s_part[3][1][e]);
      const int ch = p.stat_fold > 0 ? (col0 + e) % p.stat_fold : col0 + e;
      atomicAdd(p.stat_sum + ch, static_cast<double>(s1));
      if (p.stat_sq != nullptr) atomicAdd(p.stat_sq + ch, static_cast<double>(s2));
    }
  }
}

// ================================================================================================
// Staged-epilogue variant for narrow column tiles (BLOCK_N = 32 / 64: the layers with the most pixels, where the classic epilogue is bound by
// its own instruction count -- ncu: ~5 500 warp-instructions per 128 x 64 tile, half of them the shuffle butterflies of the BatchNorm statistics).
//   * the accumulator chunk is packed to 16 bit and written ONCE to a swizzled shared-memory tile; one elected thread hands the tile to the TMA
//     store unit (cp.async.bulk.tensor, coalesced full-line writes, hardware clipping of partial tiles);
//   * the per-channel sums of BatchNorm (sum z, sum z^2) are computed by the TENSOR CORE from the same staged tile:
//         S[., c] += ones[., 128 px] * Z[128 px, c]        (MN-major descriptor on the staged tile, as the weight-gradient kernel builds them)
//     with a second staged tile of bf16 squares for sum z^2; the two accumulators live in TMEM next to the double-buffered output accumulators
//     for the whole life of the CTA and are read once at the end.  ~100 instructions per warp and tile instead of ~700.
// Modes: EPI_F16_STATS (forward training), EPI_F16 / EPI_BF16 without addend (no statistics: staged store only).  cout % BLOCK_N == 0.
// ================================================================================================
template <int BLOCK_N, int BLOCK_K>
__global__ void __launch_bounds__(kConvThreadsP, 2)
conv_gemm_staged_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                        const __grid_constant__ ConvGemmParams p, int num_stages, int kb_per_slot, int n_tiles, int m_tiles, int out_c0) {
  using Cfg = ConvGemmCfg<BLOCK_N, BLOCK_K>;
  static_assert(BLOCK_N == 32 || BLOCK_N == 64, "staged epilogue: column tiles of 32 or 64");
  constexpr int kAccCols = BLOCK_N;                 // 32 or 64 (>= the 32-column allocation granule)
  constexpr int kTmemAlloc = 4 * kAccCols;          // two output accumulators + sum + sum of squares
  constexpr int kHalfCols = BLOCK_N / 2;            // two warps per TMEM lane quadrant, each draining half of the columns
  constexpr int CH = kHalfCols;                     // 16 or 32 columns per tcgen05.ld
  constexpr int kRowBytes = BLOCK_N * 2;            // 64 or 128: one pixel row of the staged tile = one swizzle span
  constexpr int kTileBytes = 128 * kRowBytes;
  constexpr int kEpiThreads = 256;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2 * kMaxStagesP + 5];
  __shared__ uint32_t s_tmem;
  __shared__ float s_stat[2][BLOCK_N];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  // [staged Z tile][staged squares tile][ones fp16 1 KB][ones bf16 1 KB][operand ring ...]
  const uint32_t s_z = smem_base, s_q = s_z + kTileBytes, s_one_h = s_q + kTileBytes, s_one_b = s_one_h + 1024;
  const uint32_t ring_base = s_one_b + 1024;
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kMaxStagesP]);
  const uint32_t bar_acc_full = smem_u32(&s_bar[2 * kMaxStagesP]);       // [2]
  const uint32_t bar_acc_empty = smem_u32(&s_bar[2 * kMaxStagesP + 2]);  // [2]
  const uint32_t bar_stage_free = smem_u32(&s_bar[2 * kMaxStagesP + 4]);

  const bool stats = p.epi_mode == EPI_F16_STATS;
  const bool f16 = p.epi_mode != EPI_BF16;
  const int n_tile = blockIdx.x % n_tiles;
  const int group = blockIdx.x / n_tiles;
  const int groups = gridDim.x / n_tiles;
  const int col0 = n_tile * BLOCK_N;
  const int log_tw = p.log_tw, log_th = p.log_th;
  const int num_kb = p.num_taps * p.cin_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, 8);  // one arrival per epilogue warp
    }
    mbar_init(bar_stage_free, stats ? 2u : 1u);  // the issuing thread (TMA store has read the tile) [+ tcgen05.commit of the statistics MMAs]
    mbar_fence_init();
  }
  // constant operand of the statistics MMAs: 8 rows x 128 B of 1.0 (every row group / box of the descriptor aliases these 1 KB)
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_one_h + 2 * i), "h"(static_cast<unsigned short>(0x3C00)));
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_one_b + 2 * i), "h"(static_cast<unsigned short>(0x3F80)));
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<kTmemAlloc>(smem_u32(&s_tmem));
  pdl_sync();  // the set-up above touches only this CTA's shared / tensor memory: it overlaps the tail of the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const uint32_t t_sum = tmem_base + 2 * kAccCols, t_sq = tmem_base + 3 * kAccCols;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int m = group; m < m_tiles; m += groups) {
        int t = m;
        const int tw = t % p.tiles_w;
        t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int tn = t / p.tiles_h;
        const int w0 = tw << log_tw, h0 = th << log_th, n0 = tn << (7 - log_tw - log_th);
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; kb += kb_per_slot) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, Cfg::kStageBytes * kb_per_slot);
          for (int j = 0; j < kb_per_slot; ++j) {
            const uint32_t sa = ring_base + (stage * kb_per_slot + j) * Cfg::kStageBytes;
            const ConvTap& tp = p.taps[tap];
            tma_load_5d(sa, &tmA, full, tp.c0 + cb * BLOCK_K, w0 + tp.dw, tp.p, h0 + tp.dh, n0);
            tma_load_2d(sa + Cfg::kABytes, &tmB, full, tp.kb + cb * BLOCK_K, col0);
            if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          }
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, 0);
      constexpr uint32_t lcode = umma_layout_code(Cfg::kSwizzle);
      constexpr uint32_t sbo = 8 * Cfg::kSwizzle;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int m = group; m < m_tiles; m += groups, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_acc_empty + 8 * acc, ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * kAccCols;
        for (int kb = 0; kb < num_kb; kb += kb_per_slot) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          for (int j = 0; j < kb_per_slot; ++j) {
            const uint32_t sa = ring_base + (stage * kb_per_slot + j) * Cfg::kStageBytes;
            const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t da = umma_smem_desc(sa + k * 32, 16, sbo, lcode);
              const uint64_t db = umma_smem_desc(sb + k * 32, 16, sbo, lcode);
              umma_f16(tacc, da, db, idesc, (kb | j | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_acc_full + 8 * acc);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> swizzled smem tile -> TMA store (+ statistics MMAs) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int cbeg = half * kHalfCols;
    const int row = q * 32 + lane;
    // 16-byte chunk c16 of row r lives at r * kRowBytes + ((c16 ^ f(r)) << 4): f(r) = r & 7 (128 B swizzle) or (r >> 1) & 3 (64 B swizzle)
    const uint32_t row_off = static_cast<uint32_t>(row) * kRowBytes;
    const uint32_t xr = kRowBytes == 128 ? static_cast<uint32_t>(row & 7) : static_cast<uint32_t>((row >> 1) & 3);
    const bool issuer = warp == 2 && lane == 0;
    constexpr uint32_t lcode_t = umma_layout_code(kRowBytes);
    constexpr uint32_t lcode_one = umma_layout_code(128);
    constexpr uint32_t idesc_h = umma_idesc_f16(128, BLOCK_N, 1, 1);
    constexpr uint32_t idesc_b = umma_idesc_bf16(128, BLOCK_N, 1, 1);
    int it = 0;
    for (int m = group; m < m_tiles; m += groups, ++it) {
      int t = m;
      const int tw = t % p.tiles_w;
      t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int tn = t / p.tiles_h;
      const int acc = it & 1;
      mbar_wait(bar_acc_full + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      uint32_t r[CH];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + cbeg;
      if constexpr (CH == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);  // the accumulator is in registers: the next tile's MMAs may overwrite it
      if (it > 0) mbar_wait(bar_stage_free, (it - 1) & 1);  // the previous tile has left the staging buffers
      // rows of a partial tile that lie outside the pixel grid can carry non-zero accumulators (their taps reach valid input pixels): the TMA
      // store clips them, the statistics must not see them
      const int xg = (tw << log_tw) + (row & ((1 << log_tw) - 1)), yg = (th << log_th) + ((row >> log_tw) & ((1 << log_th) - 1));
      const int ng = (tn << (7 - log_tw - log_th)) + (row >> (log_tw + log_th));
      const bool valid = xg < p.w_valid && yg < p.h_valid && ng < p.n_valid;
      uint32_t pk[CH / 2], pq[CH / 2];
#pragma unroll
      for (int i = 0; i < CH; i += 2) {
        const float a = valid ? __uint_as_float(r[i]) : 0.f, b = valid ? __uint_as_float(r[i + 1]) : 0.f;
        if (f16) {
          pk[i >> 1] = pack_f16x2(a, b);
          __half2 h;
          *reinterpret_cast<uint32_t*>(&h) = pk[i >> 1];
          const float2 f = __half22float2(h);      // squares of the STORED values
          pq[i >> 1] = pack_bf16x2(f.x * f.x, f.y * f.y);
        } else {
          pk[i >> 1] = pack_bf16x2(a, b);
        }
      }
#pragma unroll
      for (int j = 0; j < CH / 8; ++j) {
        const uint32_t c16 = static_cast<uint32_t>(cbeg / 8 + j);
        const uint32_t off = row_off + ((c16 ^ xr) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s_z + off), "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
        if (stats)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s_q + off), "r"(pq[4 * j]), "r"(pq[4 * j + 1]), "r"(pq[4 * j + 2]), "r"(pq[4 * j + 3]) : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(1, kEpiThreads);
      if (issuer) {
        tc_fence_after();
        const int w0 = tw << log_tw, h0 = th << log_th, n0 = tn << (7 - log_tw - log_th);
        tma_store_5d(&tmOut, s_z, out_c0 + col0, w0, 0, h0, n0);
        tma_store_commit();
        if (stats) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // 128 pixels = 8 K-steps of 16
            const uint64_t d1h = umma_smem_desc(s_one_h, 0, 0, lcode_one);
            const uint64_t d1b = umma_smem_desc(s_one_b, 0, 0, lcode_one);
            const uint64_t dz = umma_smem_desc(s_z + k * 16 * kRowBytes, 0, 8 * kRowBytes, lcode_t);
            const uint64_t dq = umma_smem_desc(s_q + k * 16 * kRowBytes, 0, 8 * kRowBytes, lcode_t);
            umma_f16(t_sum, d1h, dz, idesc_h, (it | k) != 0 ? 1u : 0u);
            umma_f16(t_sq, d1b, dq, idesc_b, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(bar_stage_free);
        }
        tma_store_wait_read();
        mbar_arrive(bar_stage_free);
      }
    }
    if (it > 0) mbar_wait(bar_stage_free, (it - 1) & 1);  // the last tile's statistics MMAs have completed
    tc_fence_after();
    if (issuer) tma_store_wait_all();
    if (stats && it > 0 && warp == 4) {  // warp 4 owns TMEM lanes 0..31; every row of the statistics accumulators holds the column sums
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t a[32], b[32];
        tmem_ld_32x32(t_sum + c, a);
        tmem_ld_32x32(t_sq + c, b);
        tmem_ld_wait();
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { s_stat[0][c + i] = __uint_as_float(a[i]); s_stat[1][c + i] = __uint_as_float(b[i]); }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemAlloc>(tmem_base);
  if (stats && group < m_tiles) {
    for (int e = threadIdx.x; e < BLOCK_N && col0 + e < p.cout; e += blockDim.x) {
      const int ch = p.stat_fold > 0 ? (col0 + e) % p.stat_fold : col0 + e;
      atomicAdd(p.stat_sum + ch, static_cast<double>(s_stat[0][e]));
      if (p.stat_sq != nullptr) atomicAdd(p.stat_sq + ch, static_cast<double>(s_stat[1][e]));
    }
  }
}

// ================================================================================================
// CTA-pair variant (cta_group::2) for wide layers (column tile 256): the two CTAs of a cluster compute two consecutive
// 128-pixel tiles as ONE 256 x 256 UMMA.  Each CTA loads its own activation tile and only HALF of the weight tile (the MMA
// reads B from both CTAs' shared memory), which cuts the L2->SM operand traffic per MMA by a third -- the limiter of the
// single-CTA kernel on the 3x3 layers (ncu: tensor pipe ~55 % active at 47 % L2 throughput).
//   leader CTA (cluster rank 0): issues the MMAs; its full / acc_empty barriers collect both CTAs' arrivals
//   both CTAs: TMA producer (signals the leader's full barrier), epilogue on their own TMEM half (= their own pixel tile)
// grid.x = 2 * n_tiles * groups;  pair q = blockIdx.x / 2: column tile q % n_tiles, pixel-tile pairs (q / n_tiles) + i * groups.
// ================================================================================================
template <int BLOCK_K, bool EXT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreadsP, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ ConvGemmParams p, int num_stages, int n_tiles, int m_tiles) {
  pdl_sync();
  constexpr int BLOCK_N = 256;
  constexpr int kSwizzle = BLOCK_K * 2;
  constexpr int kABytes = 128 * BLOCK_K * 2;
  constexpr int kBBytes = 128 * BLOCK_K * 2;          // this CTA's half of the 256-row weight tile
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kAccCols = 256, kTmemAlloc = 512;
  constexpr int kHalfCols = 128, CH = 32;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2 * kMaxStagesP + 4];
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[4][2][BLOCK_N];
  __shared__ __align__(16) float s_col[2][BLOCK_N];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kMaxStagesP]);
  const uint32_t bar_acc_full = smem_u32(&s_bar[2 * kMaxStagesP]);
  const uint32_t bar_acc_empty = smem_u32(&s_bar[2 * kMaxStagesP + 2]);

  const int pair = blockIdx.x >> 1;
  const int n_tile = pair % n_tiles;
  const int group = pair / n_tiles;
  const int groups = (gridDim.x >> 1) / n_tiles;
  const int col0 = n_tile * BLOCK_N;
  const int pair_tiles = (m_tiles + 1) >> 1;
  const int log_tw = p.log_tw, log_th = p.log_th;
  const int num_kb = p.num_taps * p.cin_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);    // leader: its producer's arrive.expect_tx (bytes of BOTH CTAs)
      mbar_init(bar_empty + 8 * s, 1);   // each CTA: multicast tcgen05.commit of the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, 16);  // leader: 8 epilogue warps of each CTA
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 4 * 2 * BLOCK_N; i += blockDim.x) (&s_part[0][0][0])[i] = 0.f;
  stage_col_params<BLOCK_N>(p, col0, s_col);
  if (warp == 1) tmem_alloc_pair<kTmemAlloc>(smem_u32(&s_tmem));
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // barriers of the peer are initialised before any remote arrive / TMA completion can reach them
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = group; pt < pair_tiles; pt += groups) {
        int t = 2 * pt + static_cast<int>(rank);  // a tile index >= m_tiles lands entirely out of bounds: zero fill, masked epilogue
        const int tw = t % p.tiles_w;
        t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int tn = t / p.tiles_h;
        const int w0 = tw << log_tw, h0 = th << log_th, n0 = tn << (7 - log_tw - log_th);
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t full = bar_full + 8 * stage;
          if (leader) mbar_expect_tx(full, 2 * kStageBytes);
          const ConvTap& tp = p.taps[tap];
          tma_load_5d_pair(sa, &tmA, full, tp.c0 + cb * BLOCK_K, w0 + tp.dw, tp.p, h0 + tp.dh, n0);
          tma_load_2d_pair(sa + kABytes, &tmB, full, tp.kb + cb * BLOCK_K, col0 + static_cast<int>(rank) * 128);
          if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N, 0, 0);
      constexpr uint32_t lcode = umma_layout_code(kSwizzle);
      constexpr uint32_t sbo = 8 * kSwizzle;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pt = group; pt < pair_tiles; pt += groups, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_acc_empty + 8 * acc, ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * kAccCols;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 32, 16, sbo, lcode);
            const uint64_t db = umma_smem_desc(sb + k * 32, 16, sbo, lcode);
            umma_f16_pair(tacc, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(bar_empty + 8 * stage);  // frees the slot in both CTAs
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(bar_acc_full + 8 * acc);   // both epilogues may drain
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own TMEM = own pixel tile) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int cbeg = half * kHalfCols, cend = cbeg + kHalfCols;
    // chunks that start at or beyond the last valid output channel are dead (cout not a multiple of the column tile): skip them
    const int cend_live = min(cend, ((p.cout - col0 + CH - 1) / CH) * CH);
    const bool side_is_aux = p.epi_mode == EPI_BF16_GELU_BWD || p.epi_mode == EPI_BF16_RELU_BWD;
    const __nv_bfloat16* side_base = EXT ? (side_is_aux ? p.aux_in : p.addend) : nullptr;
    const int mrow = q * 32 + lane;
    const int xl = mrow & ((1 << log_tw) - 1);
    const int yl = (mrow >> log_tw) & ((1 << log_th) - 1);
    const int nl = mrow >> (log_tw + log_th);
    int it = 0;
    for (int pt = group; pt < pair_tiles; pt += groups, ++it) {
      int t = 2 * pt + static_cast<int>(rank);
      const int tw = t % p.tiles_w;
      t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int tn = t / p.tiles_h;
      const int x = (tw << log_tw) + xl, y = (th << log_th) + yl, n = (tn << (7 - log_tw - log_th)) + nl;
      const bool valid = (x < p.w_valid) && (y < p.h_valid) && (n < p.n_valid);
      const long long pix_off = (long long)n * p.out_sn + (long long)(y * p.out_mh + p.out_ph) * p.out_sh +
                                (long long)(x * p.out_mw + p.out_pw) * p.out_sw;
      const long long add_off = (long long)n * p.add_sn + (long long)(y * p.out_mh + p.out_ph) * p.add_sh +
                                (long long)(x * p.out_mw + p.out_pw) * p.add_sw;
      const int acc = it & 1;
      const __nv_bfloat16* side_row = side_base ? side_base + (side_is_aux ? pix_off : add_off) : nullptr;
      uint4 side[CH / 8];
      if (side_base != nullptr && cend_live > cbeg) load_side_chunk<CH>(side_row, valid, col0 + cbeg, p.cout, side);  // in flight during the barrier wait
      mbar_wait(bar_acc_full + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      if (cend_live <= cbeg) {
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar_acc_empty + 8 * acc);
        continue;
      }
#pragma unroll 1
      for (int c = cbeg; c < cend_live; c += CH) {
        uint32_t r[CH];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + c, r);
        uint4 side_next[CH / 8];
        const bool more = side_base != nullptr && c + CH < cend_live;
        if (more) load_side_chunk<CH>(side_row, valid, col0 + c + CH, p.cout, side_next);
        tmem_ld_wait();
        if (c + CH >= cend_live) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(bar_acc_empty + 8 * acc);
        }
        float v[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]);
        const int cbase = col0 + c;
        conv_epilogue_chunk<CH, EXT>(p, v, valid, pix_off, add_off, cbase, lane, &s_part[q][0][c], &s_part[q][1][c], true, &s_col[0][c], &s_col[1][c],
                                side_base ? side : nullptr);
        if (more) {
#pragma unroll
          for (int k = 0; k < CH / 8; ++k) side[k] = side_next[k];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // neither CTA may free TMEM / exit while the peer's MMAs or remote arrives can still target it
  if (warp == 1) tmem_dealloc_pair<kTmemAlloc>(tmem_base);
  if (EXT ? epi_has_stats(p.epi_mode, p.stat_sum) : p.epi_mode == EPI_F16_STATS) {
    for (int e = threadIdx.x; e < BLOCK_N && col0 + e < p.cout; e += blockDim.x) {
      const float s1 = (s_part[0][0][e] + s_part[1][0][e]) + (s_part[2][0][e] + s_part[3][0][e]);
      const float s2 = (s_part[0][1][e] + s_part[1][1][e]) + (s_part[2][1][e] + s_part[3][1][e]);
      atomicAdd(p.stat_sum + col0 + e, static_cast<double>(s1));
      if (p.stat_sq != nullptr) atomicAdd(p.stat_sq + col0 + e, static_cast<double>(s2));
    }
  }
}

}  // namespace yb

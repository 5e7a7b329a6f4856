// Multi-head attention core on tcgen05: the part of nn.MultiheadAttention between in_proj and out_proj
//   out[b, i, h] = softmax_j( scale * <q[b,i,h], k[b,j,h]> + key_padding_mask[b,j] ) . v[b,j,h]
// as used by TransformerEncoderLayer / TransformerDecoderLayer (yolov7/modeling/backbone/detr_backbone.py:140,157-161,200-236: d_model 256,
// 8 heads x 32, sequences of 1050 tokens at 800x1333).  Head dimension 32 is compiled in.
//
// One CTA per (128-query tile, head, image); flash-attention style streaming over 128-key tiles:
//   warp 0     TMA producer: Q tile once, K / V tiles double-buffered (5-D NHWC maps, out-of-range tokens zero-filled)
//   warp 1     one thread issues S = Q K^T (UMMA 128x128x32, K-major operands) and O_j = P V (UMMA 128x32x128, P from shared memory,
//              V as MN-major B operand straight from its token-major tile) into TMEM
//   warps 2-5  online softmax, one query row per thread: tcgen05.ld of S, running max / sum in the exp2 domain, P written as bf16 into a
//              128B-swizzled K-major tile, running output kept in registers (acc = acc * alpha + O_j)
// With 32-wide heads the kernel is bound by MUFU.EX2 (128 x 128 exponentials per tile against 2 MFLOP of MMA), so the design goal is
// simply to keep the exponentials flowing: TMEM holds S (128 columns) and O_j (32 columns); 256 columns are allocated so that two CTAs
// share an SM and overlap each other's MMA / softmax phases.
#include "host_common.cuh"
#include "sm100.cuh"

using namespace yb;

namespace {

constexpr int kAttD = 32;          // head dimension
constexpr int kAttTile = 128;      // queries per CTA, keys per step
constexpr int kAttThreads = 192;
constexpr int kQBytes = kAttTile * kAttD * 2;       // 8 KB, 64-byte rows (swizzle 64)
constexpr int kKVBytes = kAttTile * kAttD * 2;
constexpr int kPBytes = kAttTile * kAttTile * 2;    // 32 KB: two K-blocks of [128 rows][64 keys] with 128-byte rows (swizzle 128)
constexpr int kAttSmem = kQBytes + 4 * kKVBytes + kPBytes + 1024;

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Dropout (nn.MultiheadAttention(dropout=p) on the attention probabilities, nn.Dropout on the residual branches / the FFN, detr_backbone.py:140-152,
// 200-214): a counter-based hash instead of torch's Philox stream -- keep(element) = (mix32(seed-derived key ^ index * golden) >> 8) >= p * 2^24 -- so the
// backward kernels regenerate the mask of the forward from (seed, indices) alone.  thr24 == 0 disables it (p = 0 / eval: not a single extra instruction).
struct DropParams {
  uint32_t seed, thr24;
  float inv_keep;
};
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// per-row key of the attention mask: (seed, image * heads + head, query index)
__device__ __forceinline__ uint32_t drop_row_key(uint32_t seed, uint32_t bh, uint32_t q) { return mix32(seed ^ mix32(bh * 0x9E3779B1u + q + 0x7F4A7C15u)); }
__device__ __forceinline__ float drop_factor(uint32_t row_key, uint32_t col, const DropParams& d) {
  return (mix32(row_key ^ (col * 0x9E3779B1u)) >> 8) >= d.thr24 ? d.inv_keep : 0.f;
}

struct AttnParams {
  DropParams drop;
  int lq, lk, heads;
  int q_coff, k_coff, v_coff;   // channel offsets of head 0 inside the q / k / v buffers
  float scale_log2;             // softmax scale * log2(e)
  const uint8_t* mask;          // [B][lk], 1 = ignore, may be null
  __nv_bfloat16* out;           // [B][lq][out_pitch], head h at channel out_coff + 32 h
  int out_pitch, out_coff;
  float* lse;                   // [B][heads][lq] natural-log sum-exp of the scaled scores, may be null
};

template <bool DROP>  // the dropout-free instantiation is the kernel as it was (the extra integer work and registers cost the p = 0 path 35 % when the
                      // choice was a run-time branch inside the softmax loop)
__global__ void __launch_bounds__(kAttThreads)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ AttnParams p) {
  pdl_sync();
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[8];  // q_full, kv_full[2], kv_empty[2], s_full, p_ready, o_full
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[2][kAttTile];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttTile, h = blockIdx.y, b = blockIdx.z;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + kQBytes, sV = sK + 2 * kKVBytes, sP = sV + 2 * kKVBytes;
  const uint32_t bar_q = smem_u32(&s_bar[0]), bar_kv_full = smem_u32(&s_bar[1]), bar_kv_empty = smem_u32(&s_bar[3]);
  const uint32_t bar_s = smem_u32(&s_bar[5]), bar_p = smem_u32(&s_bar[6]), bar_o = smem_u32(&s_bar[7]);
  const int ntiles = (p.lk + kAttTile - 1) / kAttTile;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_kv_full + 8 * s, 1);
      mbar_init(bar_kv_empty + 8 * s, 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, kAttTile);  // every softmax thread arrives
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(&s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = s_tmem, tmem_o = s_tmem + 128;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(bar_q, kQBytes);
      tma_load_5d(sQ, &tmQ, bar_q, p.q_coff + h * kAttD, q0, 0, 0, b);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar_kv_empty + 8 * st, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(bar_kv_full + 8 * st, 2 * kKVBytes);
        tma_load_5d(sK + st * kKVBytes, &tmK, bar_kv_full + 8 * st, p.k_coff + h * kAttD, j * kAttTile, 0, 0, b);
        tma_load_5d(sV + st * kKVBytes, &tmV, bar_kv_full + 8 * st, p.v_coff + h * kAttD, j * kAttTile, 0, 0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, kAttD, 0, 1);  // B = V tile, MN-major (head dimension contiguous)
      const uint32_t l64 = umma_layout_code(64), l128 = umma_layout_code(128);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar_kv_full + 8 * st, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAttD / 16; ++k)
          umma_f16(tmem_s, umma_smem_desc(sQ + k * 32, 16, 512, l64), umma_smem_desc(sK + st * kKVBytes + k * 32, 16, 512, l64), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_p, j & 1);  // P_j is in shared memory (and the softmax threads are done with S_j and O_{j-1})
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kAttTile / 16; ++kk) {
          const uint64_t da = umma_smem_desc(sP + (kk >> 2) * (kAttTile * 128) + (kk & 3) * 32, 16, 1024, l128);
          const uint64_t db = umma_smem_desc(sV + st * kKVBytes + kk * 16 * (kAttD * 2), kKVBytes, 8 * (kAttD * 2), l64);
          umma_f16(tmem_o, da, db, idesc_o, kk != 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
        umma_commit(bar_kv_empty + 8 * st);
      }
    }
  } else {
    const int quad = warp & 3;            // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;     // query row of this thread inside the tile
    const int tid = threadIdx.x - 64;     // 0..127 among the softmax threads
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint32_t drop_key = 0;
    if constexpr (DROP) drop_key = drop_row_key(p.drop.seed, static_cast<uint32_t>(b * p.heads + h), static_cast<uint32_t>(q0 + row));
    float acc[kAttD];
#pragma unroll
    for (int i = 0; i < kAttD; ++i) acc[i] = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      {  // additive mask of this key tile: 0 or -inf (padding keys and keys beyond lk)
        const int key = j * kAttTile + tid;
        const bool dead = key >= p.lk || (p.mask != nullptr && p.mask[static_cast<size_t>(b) * p.lk + key] != 0);
        s_bias[j & 1][tid] = dead ? -INFINITY : 0.f;
      }
      named_bar_sync(1, kAttTile);
      const float* bias = s_bias[j & 1];
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      // pass 1: row maximum of the scaled, masked scores
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < kAttTile; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_s + lane_base + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {  // 16-byte broadcast loads of the mask bias: per-element LDS made the LSU the busiest pipe
          const float4 bb = *reinterpret_cast<const float4*>(bias + c + i);
          mx = fmaxf(mx, fmaf(__uint_as_float(r[i]), p.scale_log2, bb.x));
          mx = fmaxf(mx, fmaf(__uint_as_float(r[i + 1]), p.scale_log2, bb.y));
          mx = fmaxf(mx, fmaf(__uint_as_float(r[i + 2]), p.scale_log2, bb.z));
          mx = fmaxf(mx, fmaf(__uint_as_float(r[i + 3]), p.scale_log2, bb.w));
        }
      }
      const float m_safe = mx == -INFINITY ? 0.f : mx;  // every key so far is masked: keep everything at zero without NaNs
      const float alpha = ex2(m - m_safe);               // m = -inf -> 0
      if (j > 0) {  // fold the previous tile's P V product in before P / O are overwritten
        mbar_wait(bar_o, (j - 1) & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32(tmem_o + lane_base, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < kAttD; ++i) acc[i] = (acc[i] + __uint_as_float(o[i])) * alpha;
      }
      l *= alpha;
      m = mx;
      // pass 2: probabilities -> bf16 P tile (K-major, 128-byte rows, 16-byte chunks XOR-swizzled with the row index)
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < kAttTile; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_s + lane_base + c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + c + i);
          const float p0 = ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, bb.x - m_safe));
          const float p1 = ex2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, bb.y - m_safe));
          const float p2 = ex2(fmaf(__uint_as_float(r[i + 2]), p.scale_log2, bb.z - m_safe));
          const float p3 = ex2(fmaf(__uint_as_float(r[i + 3]), p.scale_log2, bb.w - m_safe));
          rowsum += (p0 + p1) + (p2 + p3);  // the normaliser is the sum of the UNDROPPED probabilities: dropout(softmax(S)) V
          if constexpr (DROP) {
            const uint32_t col = static_cast<uint32_t>(j * kAttTile + c + i);
            pk[i >> 1] = pack_bf16x2(p0 * drop_factor(drop_key, col, p.drop), p1 * drop_factor(drop_key, col + 1, p.drop));
            pk[(i >> 1) + 1] = pack_bf16x2(p2 * drop_factor(drop_key, col + 2, p.drop), p3 * drop_factor(drop_key, col + 3, p.drop));
          } else {
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          }
        }
        const uint32_t blk = sP + (c >> 6) * (kAttTile * 128) + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 32) >> 3) + q;  // 16-byte chunk inside the 128-byte row
          const uint32_t addr = blk + (((chunk ^ (row & 7))) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                       : "memory");
        }
      }
      l += rowsum;
      fence_proxy_async();   // generic-proxy writes of P must be visible to the tensor core (async proxy)
      tc_fence_before();     // and this thread's TMEM reads are ordered before the MMA that overwrites S / O
      mbar_arrive(bar_p);
    }
    // last P V product
    mbar_wait(bar_o, (ntiles - 1) & 1);
    tc_fence_after();
    {
      uint32_t o[32];
      tmem_ld_32x32(tmem_o + lane_base, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < kAttD; ++i) acc[i] += __uint_as_float(o[i]);
    }
    const int qi = q0 + row;
    if (qi < p.lq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;  // a fully masked row gives zeros (torch gives NaN)
      __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.lq + qi) * p.out_pitch + p.out_coff + h * kAttD;
#pragma unroll
      for (int i = 0; i < kAttD; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(acc[i] * inv, acc[i + 1] * inv);
        u.y = pack_bf16x2(acc[i + 2] * inv, acc[i + 3] * inv);
        u.z = pack_bf16x2(acc[i + 4] * inv, acc[i + 5] * inv);
        u.w = pack_bf16x2(acc[i + 6] * inv, acc[i + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
      if (p.lse != nullptr) p.lse[(static_cast<size_t>(b) * p.heads + h) * p.lq + qi] = l > 0.f ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(s_tmem);
}

// ================================================================================================================================
// Backward of the attention core.  With P = softmax(S), S = scale * Q K^T + mask, O = P V and D_i = <dO_i, O_i>:
//   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - D) * scale,   dQ = dS K,   dK = dS^T Q
// P is recomputed from the saved log-sum-exp (fp32 per query row).  Two kernels, no atomics (bit-reproducible):
//   attention_bwd_kv_kernel  one CTA per (128-key tile, head, image), streams over query tiles, dK / dV accumulate in TMEM
//   attention_bwd_q_kernel   one CTA per (128-query tile, head, image), streams over key tiles, dQ accumulates in TMEM
// Operand forms (all already used by the forward kernel or the weight-gradient GEMM): S and dP are K-major x K-major UMMAs on the TMA tiles;
// P / dS are written by the softmax threads as bf16 [query][key] tiles (128-byte rows, swizzle 128) and consumed either K-major (dQ = dS K)
// or MN-major (dV = P^T dO, dK = dS^T Q: the contraction index is the tile row); dO / Q / K enter those products as MN-major B operands
// straight from their token-major tiles.
// ================================================================================================================================
struct AttnBwdParams {
  DropParams drop;
  int lq, lk, heads;
  int q_coff, k_coff, v_coff, do_coff;
  float scale, scale_log2;
  const uint8_t* mask;
  const float* lse;    // [B][heads][lq], natural log
  const float* dsum;   // [B][heads][lq], D = <dO, O>
  __nv_bfloat16 *dq, *dk, *dv;
  int dq_pitch, dq_coff, dk_pitch, dk_coff, dv_pitch, dv_coff;
};

constexpr int kBwdSmemKV = 2 * kKVBytes + 4 * kQBytes + 2 * kPBytes + 1024;  // K, V | Q x2, dO x2 | P, dS
constexpr int kBwdSmemQ = 2 * kQBytes + 4 * kKVBytes + kPBytes + 1024;       // Q, dO | K x2, V x2 | dS

// D[b][h][q] = sum_d dO[b][q][h][d] * O[b][q][h][d]
__global__ void attention_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, int o_pitch, int o_coff, const __nv_bfloat16* __restrict__ d_o, int do_pitch,
                                          int do_coff, int batch, int lq, int heads, float* __restrict__ dsum) {
  pdl_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(batch) * lq * heads) return;
  const int h = static_cast<int>(i % heads);
  const long long t = i / heads;  // b * lq + q
  const int q = static_cast<int>(t % lq), b = static_cast<int>(t / lq);
  const __nv_bfloat16* po = o + t * o_pitch + o_coff + h * kAttD;
  const __nv_bfloat16* pd = d_o + t * do_pitch + do_coff + h * kAttD;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < kAttD; c += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(po + c), g = *reinterpret_cast<const uint4*>(pd + c);
    acc += bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) + bf16_hi(a.y) * bf16_hi(g.y) +
           bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) + bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
  }
  dsum[(static_cast<size_t>(b) * heads + h) * lq + q] = acc;
}

// one 32-column chunk of P and dS for the thread's query row: reads S and dP from TMEM, writes both bf16 tiles ([query][key], swizzle 128)
// With dropout D (0 or 1/(1-p) per element): O = (D o P) V, so the tile written for dV = (D o P)^T dO is the dropped one, dP = D o (dO V^T) and
// dS = P o (dP - <dO, O>) -- the row term <dO, O> (attention_bwd_prep_kernel) already contains D through O.
template <bool DROP>
__device__ __forceinline__ void bwd_chunk(uint32_t tmem_s, uint32_t tmem_dp, uint32_t lane_base, int c, int row, const float* bias, float lse_log2, float dsum,
                                          float scale, float scale_log2, uint32_t sP, uint32_t sDS, bool write_p, const DropParams& drop, uint32_t drop_key,
                                          int key0) {
  uint32_t r[32], g[32];
  tmem_ld_32x32(tmem_s + lane_base + c, r);
  tmem_ld_32x32(tmem_dp + lane_base + c, g);
  tmem_ld_wait();
  uint32_t pk[16], dk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + c + i);
    const float p0 = ex2(fmaf(__uint_as_float(r[i]), scale_log2, bb.x - lse_log2));
    const float p1 = ex2(fmaf(__uint_as_float(r[i + 1]), scale_log2, bb.y - lse_log2));
    const float p2 = ex2(fmaf(__uint_as_float(r[i + 2]), scale_log2, bb.z - lse_log2));
    const float p3 = ex2(fmaf(__uint_as_float(r[i + 3]), scale_log2, bb.w - lse_log2));
    float f0 = 1.f, f1 = 1.f, f2 = 1.f, f3 = 1.f;
    if constexpr (DROP) {
      const uint32_t col = static_cast<uint32_t>(key0 + c + i);
      f0 = drop_factor(drop_key, col, drop); f1 = drop_factor(drop_key, col + 1, drop);
      f2 = drop_factor(drop_key, col + 2, drop); f3 = drop_factor(drop_key, col + 3, drop);
    }
    pk[i >> 1] = pack_bf16x2(p0 * f0, p1 * f1);
    pk[(i >> 1) + 1] = pack_bf16x2(p2 * f2, p3 * f3);
    dk[i >> 1] = pack_bf16x2(p0 * (__uint_as_float(g[i]) * f0 - dsum) * scale, p1 * (__uint_as_float(g[i + 1]) * f1 - dsum) * scale);
    dk[(i >> 1) + 1] = pack_bf16x2(p2 * (__uint_as_float(g[i + 2]) * f2 - dsum) * scale, p3 * (__uint_as_float(g[i + 3]) * f3 - dsum) * scale);
  }
  const uint32_t off = (c >> 6) * (kAttTile * 128) + row * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((c & 32) >> 3) + q;
    const uint32_t sw = off + (((chunk ^ (row & 7))) << 4);
    if (write_p)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + sw), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + sw), "r"(dk[4 * q]), "r"(dk[4 * q + 1]), "r"(dk[4 * q + 2]), "r"(dk[4 * q + 3]) : "memory");
  }
}

__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const uint32_t (&o)[32]) {
#pragma unroll
  for (int i = 0; i < kAttD; i += 8) {
    uint4 u;
    u.x = pack_bf16x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1]));
    u.y = pack_bf16x2(__uint_as_float(o[i + 2]), __uint_as_float(o[i + 3]));
    u.z = pack_bf16x2(__uint_as_float(o[i + 4]), __uint_as_float(o[i + 5]));
    u.w = pack_bf16x2(__uint_as_float(o[i + 6]), __uint_as_float(o[i + 7]));
    *reinterpret_cast<uint4*>(dst + i) = u;
  }
}

template <bool DROP>
__global__ void __launch_bounds__(kAttThreads)
attention_bwd_kv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ AttnBwdParams p) {
  pdl_sync();
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[8];  // kv_full, qdo_full[2], qdo_empty[2], sdp_full, pds_ready, mma_done
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[kAttTile];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kAttTile, h = blockIdx.y, b = blockIdx.z;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + kKVBytes, sQ = sV + kKVBytes, sDO = sQ + 2 * kQBytes, sP = sDO + 2 * kQBytes, sDS = sP + kPBytes;
  const uint32_t bar_kv = smem_u32(&s_bar[0]), bar_full = smem_u32(&s_bar[1]), bar_empty = smem_u32(&s_bar[3]);
  const uint32_t bar_sdp = smem_u32(&s_bar[5]), bar_pds = smem_u32(&s_bar[6]), bar_done = smem_u32(&s_bar[7]);
  const int ntiles = (p.lq + kAttTile - 1) / kAttTile;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, kAttTile);
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (threadIdx.x >= 64) {  // additive mask of this CTA's key tile
    const int tid = threadIdx.x - 64, key = k0 + tid;
    const bool dead = key >= p.lk || (p.mask != nullptr && p.mask[static_cast<size_t>(b) * p.lk + key] != 0);
    s_bias[tid] = dead ? -INFINITY : 0.f;
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = s_tmem, tmem_dp = s_tmem + 128, tmem_dv = s_tmem + 256, tmem_dk = s_tmem + 288;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_kv, 2 * kKVBytes);
      tma_load_5d(sK, &tmK, bar_kv, p.k_coff + h * kAttD, k0, 0, 0, b);
      tma_load_5d(sV, &tmV, bar_kv, p.v_coff + h * kAttD, k0, 0, 0, b);
      for (int i = 0; i < ntiles; ++i) {
        const int st = i & 1;
        mbar_wait(bar_empty + 8 * st, ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(bar_full + 8 * st, 2 * kQBytes);
        tma_load_5d(sQ + st * kQBytes, &tmQ, bar_full + 8 * st, p.q_coff + h * kAttD, i * kAttTile, 0, 0, b);
        tma_load_5d(sDO + st * kQBytes, &tmDO, bar_full + 8 * st, p.do_coff + h * kAttD, i * kAttTile, 0, 0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_t = umma_idesc_bf16(128, kAttD, 1, 1);  // A = P^T / dS^T (MN-major), B = dO / Q tile (MN-major)
      const uint32_t l64 = umma_layout_code(64), l128 = umma_layout_code(128);
      mbar_wait(bar_kv, 0);
      for (int i = 0; i < ntiles; ++i) {
        const int st = i & 1;
        mbar_wait(bar_full + 8 * st, (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAttD / 16; ++k) {
          umma_f16(tmem_s, umma_smem_desc(sQ + st * kQBytes + k * 32, 16, 512, l64), umma_smem_desc(sK + k * 32, 16, 512, l64), idesc_s, k != 0 ? 1u : 0u);
          umma_f16(tmem_dp, umma_smem_desc(sDO + st * kQBytes + k * 32, 16, 512, l64), umma_smem_desc(sV + k * 32, 16, 512, l64), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(bar_sdp);
        mbar_wait(bar_pds, i & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kAttTile / 16; ++kk) {  // contraction over the 128 query rows of the tile, 16 at a time
          const uint64_t a_p = umma_smem_desc(sP + kk * 16 * 128, kAttTile * 128, 8 * 128, l128);
          const uint64_t a_ds = umma_smem_desc(sDS + kk * 16 * 128, kAttTile * 128, 8 * 128, l128);
          const uint64_t b_do = umma_smem_desc(sDO + st * kQBytes + kk * 16 * (kAttD * 2), kQBytes, 8 * (kAttD * 2), l64);
          const uint64_t b_q = umma_smem_desc(sQ + st * kQBytes + kk * 16 * (kAttD * 2), kQBytes, 8 * (kAttD * 2), l64);
          umma_f16(tmem_dv, a_p, b_do, idesc_t, (i | kk) != 0 ? 1u : 0u);
          umma_f16(tmem_dk, a_ds, b_q, idesc_t, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(bar_done);
        umma_commit(bar_empty + 8 * st);
      }
    }
  } else {
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    for (int i = 0; i < ntiles; ++i) {
      const int qi = i * kAttTile + row;
      const size_t stat = (static_cast<size_t>(b) * p.heads + h) * p.lq + qi;
      float lse_log2 = qi < p.lq ? p.lse[stat] * 1.4426950408889634f : INFINITY;  // rows beyond lq: p = 0
      if (lse_log2 == -INFINITY) lse_log2 = INFINITY;                                 // fully masked query: zero gradients instead of NaN
      const float dsum = qi < p.lq ? p.dsum[stat] : 0.f;
      mbar_wait(bar_sdp, i & 1);
      tc_fence_after();
      if (i > 0) mbar_wait(bar_done, (i - 1) & 1);  // the previous tile's P / dS have been consumed
      uint32_t drop_key = 0;
      if constexpr (DROP) drop_key = drop_row_key(p.drop.seed, static_cast<uint32_t>(b * p.heads + h), static_cast<uint32_t>(qi));
#pragma unroll 1
      for (int c = 0; c < kAttTile; c += 32)
        bwd_chunk<DROP>(tmem_s, tmem_dp, lane_base, c, row, s_bias, lse_log2, dsum, p.scale, p.scale_log2, sP, sDS, true, p.drop, drop_key, k0);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_pds);
    }
    mbar_wait(bar_done, (ntiles - 1) & 1);
    tc_fence_after();
    const int key = k0 + row;  // accumulator row = key index
    uint32_t o[32];
    tmem_ld_32x32(tmem_dv + lane_base, o);
    tmem_ld_wait();
    if (key < p.lk) store_row32(p.dv + (static_cast<size_t>(b) * p.lk + key) * p.dv_pitch + p.dv_coff + h * kAttD, o);
    tmem_ld_32x32(tmem_dk + lane_base, o);
    tmem_ld_wait();
    if (key < p.lk) store_row32(p.dk + (static_cast<size_t>(b) * p.lk + key) * p.dk_pitch + p.dk_coff + h * kAttD, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(s_tmem);
}

template <bool DROP>
__global__ void __launch_bounds__(kAttThreads)
attention_bwd_q_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ AttnBwdParams p) {
  pdl_sync();
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[8];  // qdo_full, kv_full[2], kv_empty[2], sdp_full, ds_ready, mma_done
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[2][kAttTile];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttTile, h = blockIdx.y, b = blockIdx.z;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t sQ = base, sDO = sQ + kQBytes, sK = sDO + kQBytes, sV = sK + 2 * kKVBytes, sDS = sV + 2 * kKVBytes;
  const uint32_t bar_q = smem_u32(&s_bar[0]), bar_full = smem_u32(&s_bar[1]), bar_empty = smem_u32(&s_bar[3]);
  const uint32_t bar_sdp = smem_u32(&s_bar[5]), bar_ds = smem_u32(&s_bar[6]), bar_done = smem_u32(&s_bar[7]);
  const int ntiles = (p.lk + kAttTile - 1) / kAttTile;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_sdp, 1);
    mbar_init(bar_ds, kAttTile);
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = s_tmem, tmem_dp = s_tmem + 128, tmem_dq = s_tmem + 256;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_q, 2 * kQBytes);
      tma_load_5d(sQ, &tmQ, bar_q, p.q_coff + h * kAttD, q0, 0, 0, b);
      tma_load_5d(sDO, &tmDO, bar_q, p.do_coff + h * kAttD, q0, 0, 0, b);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar_empty + 8 * st, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(bar_full + 8 * st, 2 * kKVBytes);
        tma_load_5d(sK + st * kKVBytes, &tmK, bar_full + 8 * st, p.k_coff + h * kAttD, j * kAttTile, 0, 0, b);
        tma_load_5d(sV + st * kKVBytes, &tmV, bar_full + 8 * st, p.v_coff + h * kAttD, j * kAttTile, 0, 0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_q = umma_idesc_bf16(128, kAttD, 0, 1);  // A = dS (K-major), B = K tile (MN-major): as P V in the forward kernel
      const uint32_t l64 = umma_layout_code(64), l128 = umma_layout_code(128);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar_full + 8 * st, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAttD / 16; ++k) {
          umma_f16(tmem_s, umma_smem_desc(sQ + k * 32, 16, 512, l64), umma_smem_desc(sK + st * kKVBytes + k * 32, 16, 512, l64), idesc_s, k != 0 ? 1u : 0u);
          umma_f16(tmem_dp, umma_smem_desc(sDO + k * 32, 16, 512, l64), umma_smem_desc(sV + st * kKVBytes + k * 32, 16, 512, l64), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(bar_sdp);
        mbar_wait(bar_ds, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kAttTile / 16; ++kk) {
          const uint64_t da = umma_smem_desc(sDS + (kk >> 2) * (kAttTile * 128) + (kk & 3) * 32, 16, 1024, l128);
          const uint64_t db = umma_smem_desc(sK + st * kKVBytes + kk * 16 * (kAttD * 2), kKVBytes, 8 * (kAttD * 2), l64);
          umma_f16(tmem_dq, da, db, idesc_q, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(bar_done);
        umma_commit(bar_empty + 8 * st);
      }
    }
  } else {
    const int quad = warp & 3, row = quad * 32 + lane, tid = threadIdx.x - 64;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const int qi = q0 + row;
    const size_t stat = (static_cast<size_t>(b) * p.heads + h) * p.lq + qi;
    float lse_log2 = qi < p.lq ? p.lse[stat] * 1.4426950408889634f : INFINITY;
    if (lse_log2 == -INFINITY) lse_log2 = INFINITY;
    const float dsum = qi < p.lq ? p.dsum[stat] : 0.f;
    uint32_t drop_key = 0;
    if constexpr (DROP) drop_key = drop_row_key(p.drop.seed, static_cast<uint32_t>(b * p.heads + h), static_cast<uint32_t>(qi));
    for (int j = 0; j < ntiles; ++j) {
      {
        const int key = j * kAttTile + tid;
        const bool dead = key >= p.lk || (p.mask != nullptr && p.mask[static_cast<size_t>(b) * p.lk + key] != 0);
        s_bias[j & 1][tid] = dead ? -INFINITY : 0.f;
      }
      named_bar_sync(1, kAttTile);
      mbar_wait(bar_sdp, j & 1);
      tc_fence_after();
      if (j > 0) mbar_wait(bar_done, (j - 1) & 1);
#pragma unroll 1
      for (int c = 0; c < kAttTile; c += 32)
        bwd_chunk<DROP>(tmem_s, tmem_dp, lane_base, c, row, s_bias[j & 1], lse_log2, dsum, p.scale, p.scale_log2, sDS, sDS, false, p.drop, drop_key, j * kAttTile);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_ds);
    }
    mbar_wait(bar_done, (ntiles - 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32(tmem_dq + lane_base, o);
    tmem_ld_wait();
    if (qi < p.lq) store_row32(p.dq + (static_cast<size_t>(b) * p.lq + qi) * p.dq_pitch + p.dq_coff + h * kAttD, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(s_tmem);
}

int check_seq(const yb200_act* a, const char* name) {
  YB_REQUIRE(a && a->ptr, YB200_ERR_INVALID, "%s: null view", name);
  YB_REQUIRE(a->n > 0 && a->h == 1 && a->w > 0 && a->c > 0, YB200_ERR_INVALID, "%s: expected a [B][1][L][E] view (got %dx%dx%dx%d)", name, a->n, a->h, a->w, a->c);
  YB_REQUIRE(a->c % kAttD == 0 && a->c_pitch % 8 == 0 && a->c_off % 8 == 0 && a->c_off + a->c <= a->c_pitch, YB200_ERR_INVALID,
             "%s: channels (c=%d pitch=%d off=%d): c must be heads x 32", name, a->c, a->c_pitch, a->c_off);
  return 0;
}

}  // namespace

static int make_drop(float p_drop, uint32_t seed, DropParams* d, const char* who) {
  YB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, YB200_ERR_INVALID, "%s: dropout probability %g outside [0, 1)", who, p_drop);
  d->seed = seed;
  d->thr24 = static_cast<uint32_t>(static_cast<double>(p_drop) * 16777216.0);
  d->inv_keep = 1.f / (1.f - p_drop);
  return 0;
}
static int attention_fwd_impl(const yb200_act* q, const yb200_act* k, const yb200_act* v, const uint8_t* key_padding_mask, float scale,
                              const yb200_act* out, float* lse, float p_drop, uint32_t seed, void* stream);
extern "C" int yb200_attention_fwd(const yb200_act* q, const yb200_act* k, const yb200_act* v, const uint8_t* key_padding_mask, float scale,
                                   const yb200_act* out, float* lse, void* stream) {
  return attention_fwd_impl(q, k, v, key_padding_mask, scale, out, lse, 0.f, 0u, stream);
}
extern "C" int yb200_attention_fwd_dropout(const yb200_act* q, const yb200_act* k, const yb200_act* v, const uint8_t* key_padding_mask, float scale,
                                           const yb200_act* out, float* lse, float p_drop, uint32_t seed, void* stream) {
  return attention_fwd_impl(q, k, v, key_padding_mask, scale, out, lse, p_drop, seed, stream);
}
static int attention_fwd_impl(const yb200_act* q, const yb200_act* k, const yb200_act* v, const uint8_t* key_padding_mask, float scale,
                              const yb200_act* out, float* lse, float p_drop, uint32_t seed, void* stream) {
  int rc;
  if ((rc = check_seq(q, "attention_fwd q"))) return rc;
  if ((rc = check_seq(k, "attention_fwd k"))) return rc;
  if ((rc = check_seq(v, "attention_fwd v"))) return rc;
  if ((rc = check_seq(out, "attention_fwd out"))) return rc;
  YB_REQUIRE(k->n == q->n && v->n == q->n && out->n == q->n && k->w == v->w && out->w == q->w && k->c == q->c && v->c == q->c && out->c == q->c,
             YB200_ERR_INVALID, "attention_fwd: shapes q %dx%dx%d k %dx%dx%d v %dx%dx%d out %dx%dx%d", q->n, q->w, q->c, k->n, k->w, k->c, v->n, v->w, v->c,
             out->n, out->w, out->c);
  CUtensorMap tmQ, tmK, tmV;
  if ((rc = make_act_map(&tmQ, *q, false, kAttD, kAttTile, 1, 1))) return rc;
  if ((rc = make_act_map(&tmK, *k, false, kAttD, kAttTile, 1, 1))) return rc;
  if ((rc = make_act_map(&tmV, *v, false, kAttD, kAttTile, 1, 1))) return rc;
  AttnParams p;
  if ((rc = make_drop(p_drop, seed, &p.drop, "attention_fwd"))) return rc;
  p.lq = q->w; p.lk = k->w; p.heads = q->c / kAttD;
  p.q_coff = q->c_off; p.k_coff = k->c_off; p.v_coff = v->c_off;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mask = key_padding_mask;
  p.out = static_cast<__nv_bfloat16*>(out->ptr);
  p.out_pitch = out->c_pitch; p.out_coff = out->c_off;
  p.lse = lse;
  static PerDevice<bool> attr_set_dev(false);
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    YB_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.lq, kAttTile), p.heads, q->n);
  if (p.drop.thr24 != 0)
    launch_k(attention_fwd_kernel<true>, grid, kAttThreads, kAttSmem, as_stream(stream), tmQ, tmK, tmV, p);
  else
    launch_k(attention_fwd_kernel<false>, grid, kAttThreads, kAttSmem, as_stream(stream), tmQ, tmK, tmV, p);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t yb200_attention_bwd_workspace(const yb200_act* q) {
  if (!q || q->n <= 0 || q->w <= 0 || q->c <= 0 || q->c % kAttD != 0) return YB200_ERR_INVALID;
  return 4LL * q->n * (q->c / kAttD) * q->w;
}

static int attention_bwd_impl(const yb200_act* q, const yb200_act* k, const yb200_act* v, const yb200_act* out, const yb200_act* dout,
                              const uint8_t* key_padding_mask, float scale, const float* lse, const yb200_act* dq, const yb200_act* dk,
                              const yb200_act* dv, void* workspace, float p_drop, uint32_t seed, void* stream);
extern "C" int yb200_attention_bwd(const yb200_act* q, const yb200_act* k, const yb200_act* v, const yb200_act* out, const yb200_act* dout,
                                   const uint8_t* key_padding_mask, float scale, const float* lse, const yb200_act* dq, const yb200_act* dk,
                                   const yb200_act* dv, void* workspace, void* stream) {
  return attention_bwd_impl(q, k, v, out, dout, key_padding_mask, scale, lse, dq, dk, dv, workspace, 0.f, 0u, stream);
}
extern "C" int yb200_attention_bwd_dropout(const yb200_act* q, const yb200_act* k, const yb200_act* v, const yb200_act* out, const yb200_act* dout,
                                           const uint8_t* key_padding_mask, float scale, const float* lse, const yb200_act* dq, const yb200_act* dk,
                                           const yb200_act* dv, void* workspace, float p_drop, uint32_t seed, void* stream) {
  return attention_bwd_impl(q, k, v, out, dout, key_padding_mask, scale, lse, dq, dk, dv, workspace, p_drop, seed, stream);
}
static int attention_bwd_impl(const yb200_act* q, const yb200_act* k, const yb200_act* v, const yb200_act* out, const yb200_act* dout,
                              const uint8_t* key_padding_mask, float scale, const float* lse, const yb200_act* dq, const yb200_act* dk,
                              const yb200_act* dv, void* workspace, float p_drop, uint32_t seed, void* stream) {
  int rc;
  const yb200_act* all[8] = {q, k, v, out, dout, dq, dk, dv};
  const char* names[8] = {"attention_bwd q", "attention_bwd k", "attention_bwd v", "attention_bwd out", "attention_bwd dout", "attention_bwd dq", "attention_bwd dk",
                          "attention_bwd dv"};
  for (int i = 0; i < 8; ++i)
    if ((rc = check_seq(all[i], names[i]))) return rc;
  YB_REQUIRE(lse && workspace, YB200_ERR_INVALID, "attention_bwd: null lse / workspace");
  const yb200_act* qlike[3] = {out, dout, dq};
  for (const yb200_act* t : qlike) YB_REQUIRE(t->n == q->n && t->w == q->w && t->c == q->c, YB200_ERR_INVALID, "attention_bwd: query-side shapes differ");
  const yb200_act* klike[4] = {k, v, dk, dv};
  for (const yb200_act* t : klike) YB_REQUIRE(t->n == q->n && t->w == k->w && t->c == q->c, YB200_ERR_INVALID, "attention_bwd: key-side shapes differ");
  CUtensorMap tmQ, tmK, tmV, tmDO;
  if ((rc = make_act_map(&tmQ, *q, false, kAttD, kAttTile, 1, 1))) return rc;
  if ((rc = make_act_map(&tmK, *k, false, kAttD, kAttTile, 1, 1))) return rc;
  if ((rc = make_act_map(&tmV, *v, false, kAttD, kAttTile, 1, 1))) return rc;
  if ((rc = make_act_map(&tmDO, *dout, false, kAttD, kAttTile, 1, 1))) return rc;
  AttnBwdParams p;
  if ((rc = make_drop(p_drop, seed, &p.drop, "attention_bwd"))) return rc;
  p.lq = q->w; p.lk = k->w; p.heads = q->c / kAttD;
  p.q_coff = q->c_off; p.k_coff = k->c_off; p.v_coff = v->c_off; p.do_coff = dout->c_off;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.mask = key_padding_mask;
  p.lse = lse;
  p.dsum = static_cast<const float*>(workspace);
  p.dq = static_cast<__nv_bfloat16*>(dq->ptr); p.dq_pitch = dq->c_pitch; p.dq_coff = dq->c_off;
  p.dk = static_cast<__nv_bfloat16*>(dk->ptr); p.dk_pitch = dk->c_pitch; p.dk_coff = dk->c_off;
  p.dv = static_cast<__nv_bfloat16*>(dv->ptr); p.dv_pitch = dv->c_pitch; p.dv_coff = dv->c_off;
  static PerDevice<bool> attr_set_dev(false);
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemKV));
    YB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_q_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemQ));
    YB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemKV));
    YB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_q_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemQ));
    attr_set = true;
  }
  cudaStream_t st = as_stream(stream);
  const long long rows = 1LL * q->n * q->w * p.heads;
  launch_k(attention_bwd_prep_kernel, static_cast<int>((rows + 255) / 256), 256, 0, st, static_cast<const __nv_bfloat16*>(out->ptr), out->c_pitch, out->c_off,
                                                                                  static_cast<const __nv_bfloat16*>(dout->ptr), dout->c_pitch, dout->c_off, q->n,
                                                                                  q->w, p.heads, static_cast<float*>(workspace));
  if (p.drop.thr24 != 0) {
    launch_k(attention_bwd_kv_kernel<true>, dim3(ceil_div(p.lk, kAttTile), p.heads, q->n), kAttThreads, kBwdSmemKV, st, tmQ, tmK, tmV, tmDO, p);
    launch_k(attention_bwd_q_kernel<true>, dim3(ceil_div(p.lq, kAttTile), p.heads, q->n), kAttThreads, kBwdSmemQ, st, tmQ, tmK, tmV, tmDO, p);
  } else {
    launch_k(attention_bwd_kv_kernel<false>, dim3(ceil_div(p.lk, kAttTile), p.heads, q->n), kAttThreads, kBwdSmemKV, st, tmQ, tmK, tmV, tmDO, p);
    launch_k(attention_bwd_q_kernel<false>, dim3(ceil_div(p.lq, kAttTile), p.heads, q->n), kAttThreads, kBwdSmemQ, st, tmQ, tmK, tmV, tmDO, p);
  }
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------------
// nn.Dropout on a [B][1][L][C] bf16 activation (dropout / dropout1 / dropout2 / dropout3 of the transformer layers, detr_backbone.py:147-152, 207-214):
//   out = residual + x * keep(seed, element) / (1 - p) * extra_scale        (residual may be null; extra_scale = 1 except in the FFN backward)
// with the same counter-based hash as the attention kernels; `element` is the logical index ((b * L + l) * C + c), so views with different channel
// pitches share a mask and the backward pass (the same call on the gradient, same seed) regenerates it.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x, int x_pitch, const __nv_bfloat16* __restrict__ res, int res_pitch,
                                    __nv_bfloat16* __restrict__ out, int out_pitch, long long rows, int c, DropParams d, float extra_scale) {
  pdl_sync();
  const int cv = c / 8;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cv;
    const int c8 = static_cast<int>(i - r * cv) * 8;
    const uint4 xv = *reinterpret_cast<const uint4*>(x + r * x_pitch + c8);
    uint4 rv = make_uint4(0u, 0u, 0u, 0u);
    if (res) rv = *reinterpret_cast<const uint4*>(res + r * res_pitch + c8);
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, rs[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t o[4];
    const uint32_t e0 = static_cast<uint32_t>(r * c + c8);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float f0 = (mix32(d.seed ^ ((e0 + 2 * k) * 0x9E3779B1u)) >> 8) >= d.thr24 ? d.inv_keep * extra_scale : 0.f;
      const float f1 = (mix32(d.seed ^ ((e0 + 2 * k + 1) * 0x9E3779B1u)) >> 8) >= d.thr24 ? d.inv_keep * extra_scale : 0.f;
      o[k] = pack_bf16x2(fmaf(bf16_lo(xs[k]), f0, bf16_lo(rs[k])), fmaf(bf16_hi(xs[k]), f1, bf16_hi(rs[k])));
    }
    *reinterpret_cast<uint4*>(out + r * out_pitch + c8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
}  // namespace

extern "C" int yb200_dropout(const yb200_act* x, const yb200_act* residual, const yb200_act* out, float p_drop, uint32_t seed, float extra_scale,
                             void* stream) {
  int rc;
  if ((rc = check_seq(x, "dropout x"))) return rc;
  if ((rc = check_seq(out, "dropout out"))) return rc;
  if (residual && (rc = check_seq(residual, "dropout residual"))) return rc;
  YB_REQUIRE(out->n == x->n && out->w == x->w && out->c == x->c && (!residual || (residual->n == x->n && residual->w == x->w && residual->c == x->c)),
             YB200_ERR_INVALID, "dropout: shapes differ");
  YB_REQUIRE(1LL * x->n * x->w * x->c < (1LL << 32), YB200_ERR_UNSUPPORTED, "dropout: more than 2^32 elements");
  DropParams d;
  if ((rc = make_drop(p_drop, seed, &d, "dropout"))) return rc;
  const long long rows = 1LL * x->n * x->w;
  const long long total = rows * (x->c / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 16LL * sm_count()));
  launch_k(dropout_bf16_kernel, blocks, 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(x->ptr) + x->c_off, x->c_pitch,
           residual ? static_cast<const __nv_bfloat16*>(residual->ptr) + residual->c_off : nullptr, residual ? residual->c_pitch : 0,
           static_cast<__nv_bfloat16*>(out->ptr) + out->c_off, out->c_pitch, rows, x->c, d, extra_scale);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

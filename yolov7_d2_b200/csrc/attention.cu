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

struct AttnParams {
  int lq, lk, heads;
  int q_coff, k_coff, v_coff;   // channel offsets of head 0 inside the q / k / v buffers
  float scale_log2;             // softmax scale * log2(e)
  const uint8_t* mask;          // [B][lk], 1 = ignore, may be null
  __nv_bfloat16* out;           // [B][lq][out_pitch], head h at channel out_coff + 32 h
  int out_pitch, out_coff;
  float* lse;                   // [B][heads][lq] natural-log sum-exp of the scaled scores, may be null
};

__global__ void __launch_bounds__(kAttThreads)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ AttnParams p) {
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
          rowsum += (p0 + p1) + (p2 + p3);
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
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

int check_seq(const yb200_act* a, const char* name) {
  YB_REQUIRE(a && a->ptr, YB200_ERR_INVALID, "%s: null view", name);
  YB_REQUIRE(a->n > 0 && a->h == 1 && a->w > 0 && a->c > 0, YB200_ERR_INVALID, "%s: expected a [B][1][L][E] view (got %dx%dx%dx%d)", name, a->n, a->h, a->w, a->c);
  YB_REQUIRE(a->c % kAttD == 0 && a->c_pitch % 8 == 0 && a->c_off % 8 == 0 && a->c_off + a->c <= a->c_pitch, YB200_ERR_INVALID,
             "%s: channels (c=%d pitch=%d off=%d): c must be heads x 32", name, a->c, a->c_pitch, a->c_off);
  return 0;
}

}  // namespace

extern "C" int yb200_attention_fwd(const yb200_act* q, const yb200_act* k, const yb200_act* v, const uint8_t* key_padding_mask, float scale,
                                   const yb200_act* out, float* lse, void* stream) {
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
  p.lq = q->w; p.lk = k->w; p.heads = q->c / kAttD;
  p.q_coff = q->c_off; p.k_coff = k->c_off; p.v_coff = v->c_off;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mask = key_padding_mask;
  p.out = static_cast<__nv_bfloat16*>(out->ptr);
  p.out_pitch = out->c_pitch; p.out_coff = out->c_off;
  p.lse = lse;
  static bool attr_set = false;
  if (!attr_set) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.lq, kAttTile), p.heads, q->n);
  attention_fwd_kernel<<<grid, kAttThreads, kAttSmem, as_stream(stream)>>>(tmQ, tmK, tmV, p);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Attention kernels (head_dim = 64).
//   * attn_varlen_simt : softmax(q k^T / 8 + mask) v over packed ragged sequences, fp32 math in a
//     fixed order -- prefill of the AR decoder, NAR passes and the training forward in fp32
//     parity mode.  Also fills the KV cache.
//   * attn_decode      : one query row per (utterance, head) against the growing KV cache.
//     HBM-bound: 16-byte coalesced K/V reads, warp-shuffle dot products and softmax
//     reductions, split-KV across CTAs when B*H is too small to fill 148 SMs.
//
// Reference arithmetic: F.multi_head_attention_forward as called from
// valle/modules/activation.py:408-427; masks valle/models/valle.py:1010-1033 (AR) / none (NAR).
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vb {

static constexpr int HD = 64;

// ------------------------------------------------------------------------------------------
// Ragged multi-query attention, 64x64 tiles, 256 threads, 4x4 micro-tiles.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
attn_varlen_simt_kernel(const T *__restrict__ qkv, int n_head, const int32_t *__restrict__ cu_seqlens,
                        const int32_t *__restrict__ text_lens, const int32_t *__restrict__ seg1_lens,
                        int seg1_start, int mask_mode, T *__restrict__ out,
                        T *__restrict__ kcache, T *__restrict__ vcache, int64_t cache_seq_stride,
                        int cache_cap, const uint8_t *__restrict__ dmask, int64_t dld, DropCfg drop) {
  constexpr int LDT = 68;  // padded leading dim (floats), keeps float4 alignment
  extern __shared__ __align__(16) float smem[];
  float *Qt = smem;             // [64 e][LDT rows]
  float *Kt = Qt + 64 * LDT;    // [64 e][LDT keys]
  float *Vs = Kt + 64 * LDT;    // [64 keys][LDT e]
  float *Pt = Vs + 64 * LDT;    // [64 keys][LDT rows]

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = cu_seqlens[b], L = cu_seqlens[b + 1] - r0;
  const int q0 = blockIdx.x * 64;
  if (q0 >= L) return;
  const int S = (mask_mode != VB_MASK_FULL) ? text_lens[b] : 0;
  const int c1 = (mask_mode >= VB_MASK_PADDED_AR) ? seg1_lens[b] : 0;
  const int d = n_head * HD;
  const int64_t ld = 3 * (int64_t)d;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, le0 = (tid & 3) * 16;

  // load Q tile (transposed)
  {
    const int qr = q0 + lrow;
    const T *src = qkv + (int64_t)(r0 + min(qr, L - 1)) * ld + h * HD + le0;
#pragma unroll
    for (int i = 0; i < 16; ++i) Qt[(le0 + i) * LDT + lrow] = (qr < L) ? to_f32(src[i]) : 0.f;
  }
  RowMask lim[4];  // visibility rule per owned row
#pragma unroll
  for (int i = 0; i < 4; ++i) lim[i] = make_row_mask(mask_mode, q0 + ty * 4 + i, L, S, seg1_start, c1);
  const int q_hi = min(q0 + 64, L);
  const int kv_max = (mask_mode == VB_MASK_VALLE_AR) ? max(S, q_hi) : L;

  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -CUDART_INF_F;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  for (int j0 = 0; j0 < kv_max; j0 += 64) {
    __syncthreads();  // previous tile fully consumed (also covers the Q store above)
    {
      const int kr = j0 + lrow;
      const bool ok = kr < L;
      const T *ksrc = qkv + (int64_t)(r0 + min(kr, L - 1)) * ld + d + h * HD + le0;
      const T *vsrc = ksrc + d;
      T kraw[16], vraw[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        kraw[i] = ksrc[i];
        vraw[i] = vsrc[i];
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        Kt[(le0 + i) * LDT + lrow] = ok ? to_f32(kraw[i]) : 0.f;
        Vs[lrow * LDT + le0 + i] = ok ? to_f32(vraw[i]) : 0.f;
      }
      if (kcache != nullptr && j0 == q0 && ok) {  // this CTA owns rows [q0, q0+64) of the cache
        const int64_t off = (int64_t)b * cache_seq_stride + ((int64_t)h * cache_cap + kr) * HD + le0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          kcache[off + i] = kraw[i];
          vcache[off + i] = vraw[i];
        }
      }
    }
    __syncthreads();
    // S = Q K^T
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int e = 0; e < HD; ++e) {
      const float4 qa = *reinterpret_cast<const float4 *>(&Qt[e * LDT + ty * 4]);
      const float4 kb = *reinterpret_cast<const float4 *>(&Kt[e * LDT + tx * 4]);
      const float qv[4] = {qa.x, qa.y, qa.z, qa.w};
      const float kv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
    // online softmax per row (16 lanes share a row)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = j0 + tx * 4 + j;
        bool seen = lim[i].ok(c);
        if (dmask != nullptr && seen) seen = dmask[(int64_t)(q0 + ty * 4 + i) * dld + c] == 0;  // True = blocked
        s[i][j] = seen ? s[i][j] * 0.125f : -CUDART_INF_F;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float corr = expf(m_run[i] - m_use);  // exp(-inf) = 0 on the first tile
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = expf(s[i][j] - m_use);
        rs += s[i][j];
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[i] = l_run[i] * corr + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
      if (drop.thresh != 0) {   // training: dropout on the normalised probabilities = on p~ with the full row sum kept
        const uint64_t base = ((uint64_t)(b * n_head + h) * drop.lmax + (q0 + ty * 4 + i)) * drop.lmax + j0 + tx * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = drop_keep(drop, base + j) ? s[i][j] * drop.inv_keep : 0.f;
      }
    }
    // P^T to shared: Pt[key][row]
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4 *>(&Pt[(tx * 4 + j) * LDT + ty * 4]) =
          make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
    // O += P V
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      const float4 pa = *reinterpret_cast<const float4 *>(&Pt[c * LDT + ty * 4]);
      const float4 vb4 = *reinterpret_cast<const float4 *>(&Vs[c * LDT + tx * 4]);
      const float pv[4] = {pa.x, pa.y, pa.z, pa.w};
      const float vv[4] = {vb4.x, vb4.y, vb4.z, vb4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pv[i], vv[j], o[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qr = q0 + ty * 4 + i;
    if (qr >= L) continue;
    const float inv = 1.f / l_run[i];
    T *dst = out + (int64_t)(r0 + qr) * d + h * HD + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = from_f32<T>(o[i][j] * inv);
  }
}

int launch_attention_varlen(const void *qkv, int dtype, int64_t M, int B, int n_head, int head_dim,
                            const int32_t *cu_seqlens, const int32_t *text_lens, const int32_t *seg1_lens,
                            int seg1_start, int max_seqlen, int mask_mode, void *out, void *kcache, void *vcache,
                            int64_t cache_seq_stride, int cache_cap, const uint8_t *dense_mask, int64_t dense_ld,
                            cudaStream_t s, const DropCfg *drop) {
  VB_CHECK_ARG(head_dim == HD, "attention: head_dim=%d, only 64 is built", head_dim);
  DropCfg dc{};
  if (drop) dc = *drop;
  dc.lmax = max_seqlen;
  const bool dropping = dc.thresh != 0;   // attention-probability dropout: the CUDA-core kernel carries the mask
  VB_CHECK_ARG(mask_mode >= VB_MASK_FULL && mask_mode <= VB_MASK_DENSE, "attention: bad mask mode %d", mask_mode);
  if (mask_mode == VB_MASK_DENSE) {
    VB_CHECK_ARG(dense_mask != nullptr && dense_ld >= max_seqlen, "attention: VB_MASK_DENSE needs a [>=L, >=L] byte mask");
    mask_mode = VB_MASK_FULL;  // every key of the sequence, minus the blocked entries of the dense mask
  } else {
    dense_mask = nullptr;
  }
  VB_CHECK_ARG(mask_mode == VB_MASK_FULL || text_lens != nullptr, "attention: this mask mode needs text_lens");
  VB_CHECK_ARG(mask_mode < VB_MASK_PADDED_AR || seg1_lens != nullptr, "attention: padded mask modes need seg1_lens");
  if (M == 0 || B == 0) return VB_OK;
  const size_t smem = 4 * 64 * 68 * sizeof(float);
  dim3 grid((max_seqlen + 63) / 64, n_head, B);
  if (dtype == VB_F32) {
    auto k = attn_varlen_simt_kernel<float>;
    VB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>((const float *)qkv, n_head, cu_seqlens, text_lens, seg1_lens, seg1_start, mask_mode, (float *)out,
                              (float *)kcache, (float *)vcache, cache_seq_stride, cache_cap, dense_mask, dense_ld, dc);
  } else if (dtype == VB_BF16 && !dropping && dense_mask == nullptr && kcache == nullptr && getenv("VB_ATTN_SIMT") == nullptr &&
             attention_tcgen05_enabled()) {
    // 128-row query tiles on tcgen05/TMEM; a ragged last tile is shifted back to [L-128, L) (overlap, rows are
    // independent) so that a length like 1025 costs 9 tiles, not 9 tiles plus a 64-row warp-level pass
    return launch_attention_tcgen05((const bf16 *)qkv, M, B, n_head, cu_seqlens, text_lens, seg1_lens, seg1_start,
                                    max_seqlen, mask_mode, (bf16 *)out, 2, s);
  } else if (dtype == VB_BF16 && !dropping && dense_mask == nullptr && getenv("VB_ATTN_SIMT") == nullptr) {
    return launch_attention_mma((const bf16 *)qkv, M, B, n_head, cu_seqlens, text_lens, seg1_lens, seg1_start, max_seqlen, mask_mode,
                                (bf16 *)out, (bf16 *)kcache, (bf16 *)vcache, cache_seq_stride, cache_cap, 0, s);
  } else if (dtype == VB_BF16) {
    auto k = attn_varlen_simt_kernel<bf16>;
    VB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>((const bf16 *)qkv, n_head, cu_seqlens, text_lens, seg1_lens, seg1_start, mask_mode, (bf16 *)out,
                              (bf16 *)kcache, (bf16 *)vcache, cache_seq_stride, cache_cap, dense_mask, dense_ld, dc);
  } else {
    set_error("attention: bad dtype %d", dtype);
    return VB_ERR_ARG;
  }
  VB_LAUNCH_CHECK();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------
// Single-query decode attention over the KV cache.
//   grid (H, B, nsplit), 128 threads.  kv_len[b] = S_b + Tp_b + n_gen[b].
// ------------------------------------------------------------------------------------------
static constexpr int kDecMaxChunk = 4096;

template <typename T> struct KvRow8 {  // 8 consecutive elements of a cache row as floats
  static __device__ __forceinline__ void load(const T *p, float (&f)[8]);
};
template <> __device__ __forceinline__ void KvRow8<float>::load(const float *p, float (&f)[8]) {
  const uint4 a = ldg_stream16(p), b = ldg_stream16(p + 4);
  f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
  f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
}
template <> __device__ __forceinline__ void KvRow8<bf16>::load(const bf16 *p, float (&f)[8]) {
  Vec16<bf16> v;
  v.raw = ldg_stream16(p);
  v.unpack(f);
}

// Optional fused prologue: q/k/v of the CURRENT token arrive as split-K partials of the QKV
// projection (gemm_decode.cu); the CTA sums them in fixed order, adds the bias, appends k/v to the
// cache (split 0) and serves the new key/value from shared memory.
struct QkvPartials {
  const float *part;  // [splits][64][ldp] or nullptr
  const float *bias;  // [3d] (with a folded LayerNorm: bias + beta W^T)
  int splits, ldp;
  LnFoldStats fold;   // fold.stats != NULL: the partials are x (gamma o W)^T of the raw rows (gemm_decode_x_kernel)
};

// A finished utterance (stop rule fired, vb_ar_state.finished != 0) takes no further part in the step: its KV
// cache stays as it is (nothing appended, nothing streamed) and its attention output row is zero.  Uniform
// over the CTA.  Returns true if the CTA is done.
__device__ __forceinline__ bool decode_row_finished(const int32_t *finished, int b, int h, int sp, int d, int tid,
                                                    int nsplit, int n_head, float *out, bf16 *out16, float *part_o,
                                                    float *part_ml) {
  if (finished == nullptr || finished[b] == 0) return false;
  if (tid < HD) {
    if (nsplit == 1) {
      out[(int64_t)b * d + h * HD + tid] = 0.f;
      if (out16) out16[(int64_t)b * d + h * HD + tid] = __float2bfloat16_rn(0.f);
    } else {
      const int64_t pi = ((int64_t)b * n_head + h) * nsplit + sp;
      part_o[pi * HD + tid] = 0.f;
      if (tid == 0) {
        part_ml[pi * 2] = -CUDART_INF_F;
        part_ml[pi * 2 + 1] = 0.f;
      }
    }
  }
  return true;
}

// Single pass, no block-level synchronisation inside the loop: each 8-lane group owns every 16th key of
// the chunk, streams the K row and the V row of its keys together (4 keys = 8 x 16-byte loads in
// flight per lane), and keeps its OWN online-softmax state (m, l, 8 output elements per lane).  The 16
// groups are merged once at the end (flash-decoding style).
template <typename T>
__global__ void __launch_bounds__(128)
attn_decode_kernel(const float *__restrict__ q, QkvPartials qp, int n_head, T *__restrict__ kcache,
                   T *__restrict__ vcache, int64_t cache_seq_stride, int cache_cap,
                   const int32_t *__restrict__ text_len, const int32_t *__restrict__ prompt_len,
                   const int32_t *__restrict__ n_gen, const int32_t *__restrict__ finished,
                   float *__restrict__ out, bf16 *__restrict__ out16,
                   float *__restrict__ part_o, float *__restrict__ part_ml, int nsplit) {
  __shared__ __align__(16) float qs[HD];
  __shared__ __align__(16) float knew[HD];
  __shared__ __align__(16) float vnew[HD];
  __shared__ float g_o[16][HD + 1];
  __shared__ float g_m[16], g_l[16];
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = n_head * HD;
  pdl_wait();
  if (decode_row_finished(finished, b, h, sp, d, threadIdx.x, nsplit, n_head, out, out16, part_o, part_ml)) return;
  int kv_len = text_len[b] + prompt_len[b] + n_gen[b];
  kv_len = max(1, min(kv_len, cache_cap));
  const int pos = kv_len - 1;  // cache row of the current token
  const int chunk = ((kv_len + nsplit - 1) / nsplit + 15) & ~15;
  const int c0 = sp * chunk, c1 = min(kv_len, c0 + chunk);
  const int n = max(0, c1 - c0);
  T *kb = kcache + (int64_t)b * cache_seq_stride + (int64_t)h * cache_cap * HD;
  T *vb_ = vcache + (int64_t)b * cache_seq_stride + (int64_t)h * cache_cap * HD;
  const bool has_new = qp.part != nullptr;
  if (tid < HD) {
    if (has_new) {
      float a[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = j * d + h * HD + tid;
        const float *p = qp.part + (int64_t)b * qp.ldp + col;
        float acc = __ldcg(p);
#pragma unroll 6
        for (int s = 1; s < qp.splits; ++s) acc += __ldcg(p + (int64_t)s * 64 * qp.ldp);
        if (qp.fold.stats) {
          float mean, rstd;
          ln_fold_moments(qp.fold, b, h, mean, rstd);
          acc = rstd * (acc - mean * qp.fold.c[col]);
        }
        a[j] = acc + qp.bias[col];
      }
      qs[tid] = a[0] * 0.125f;
      const T k16 = from_f32<T>(a[1]), v16 = from_f32<T>(a[2]);
      knew[tid] = to_f32(k16);  // exactly what later steps will read back from the cache
      vnew[tid] = to_f32(v16);
      if (sp == 0) {
        kb[(int64_t)pos * HD + tid] = k16;
        vb_[(int64_t)pos * HD + tid] = v16;
      }
    } else {
      qs[tid] = q[(int64_t)b * d + h * HD + tid] * 0.125f;
    }
  }
  __syncthreads();

  const int grp = warp * 4 + (lane >> 3);  // 0..15
  const int j8 = (lane & 7) * 8;           // this lane's 8 head dims
  float qf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qf[i] = qs[j8 + i];
  float m = -CUDART_INF_F, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  for (int base = 0; base < n; base += 64) {
    float kf[4][8], vf[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = base + u * 16 + grp;
      const int kk = min(key, n - 1);
      KvRow8<T>::load(kb + (int64_t)(c0 + kk) * HD + j8, kf[u]);
      KvRow8<T>::load(vb_ + (int64_t)(c0 + kk) * HD + j8, vf[u]);
    }
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = base + u * 16 + grp;
      if (has_new && c0 + min(key, n - 1) == pos) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          kf[u][i] = knew[j8 + i];
          vf[u][i] = vnew[j8 + i];
        }
      }
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(qf[i], kf[u][i], dot);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      s[u] = key < n ? dot : -CUDART_INF_F;
    }
    const float mt = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    const float mn = fmaxf(m, mt);  // finite: key `base + grp` < n whenever this group has work ...
    if (mn != -CUDART_INF_F) {      // ... otherwise the group has no key in this tile at all
      const float corr = expf(m - mn);
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = expf(s[u] - mn);
      l = l * corr + ((p[0] + p[1]) + (p[2] + p[3]));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = acc[i] * corr;
#pragma unroll
        for (int u = 0; u < 4; ++u) a = fmaf(p[u], vf[u][i], a);
        acc[i] = a;
      }
      m = mn;
    }
  }
  // ---- merge the 16 groups ------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) g_o[grp][j8 + i] = acc[i];
  if ((lane & 7) == 0) {
    g_m[grp] = m;
    g_l[grp] = l;
  }
  __syncthreads();
  if (tid < HD) {
    float mm = -CUDART_INF_F;
#pragma unroll
    for (int g = 0; g < 16; ++g) mm = fmaxf(mm, g_m[g]);
    float lt = 0.f, ot = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      if (g_m[g] == -CUDART_INF_F) continue;
      const float w = expf(g_m[g] - mm);
      lt += g_l[g] * w;
      ot += g_o[g][tid] * w;
    }
    if (nsplit == 1) {
      out[(int64_t)b * d + h * HD + tid] = ot / lt;
      if (out16) out16[(int64_t)b * d + h * HD + tid] = __float2bfloat16_rn(ot / lt);
    } else {
      const int64_t pi = ((int64_t)b * n_head + h) * nsplit + sp;
      part_o[pi * HD + tid] = ot;
      if (tid == 0) {
        part_ml[pi * 2] = n > 0 ? mm : -CUDART_INF_F;
        part_ml[pi * 2 + 1] = n > 0 ? lt : 0.f;
      }
    }
  }
}

// bf16 two-phase kernel (scores of the whole chunk to shared memory, block softmax, then P.V) with the FIRST batch
// of K rows fetched ahead of the q/k/v prologue (and, with the fused QKV prologue, ahead of the dependency wait)
// and the first batch of V rows fetched ahead of the block softmax: all CTAs of the single wave run their phases
// in lock step, so without this the HBM pipe idles through every prologue / softmax / epilogue of the launch.
template <int U>
__global__ void __launch_bounds__(128, 7)
attn_decode_2phase_pf_kernel(const float *__restrict__ q, QkvPartials qp, int n_head, bf16 *__restrict__ kcache,
                   bf16 *__restrict__ vcache, int64_t cache_seq_stride, int cache_cap,
                   const int32_t *__restrict__ text_len, const int32_t *__restrict__ prompt_len,
                   const int32_t *__restrict__ n_gen, const int32_t *__restrict__ finished,
                   float *__restrict__ out, bf16 *__restrict__ out16,
                   float *__restrict__ part_o, float *__restrict__ part_ml, int nsplit) {
  // score buffer of the chunk: dynamic shared memory sized by the launch (cache_cap / nsplit keys), so that the
  // kernel's footprint -- and with it the shared-memory carve-out the driver picks, i.e. how much L1 is left to land
  // the ~110 KB of K / V loads an SM keeps in flight -- follows the actual context instead of the 4096-key maximum
  extern __shared__ float sc[];
  __shared__ __align__(16) float qs[HD];
  __shared__ __align__(16) float knew[HD];
  __shared__ __align__(16) float vnew[HD];
  __shared__ float red[16][HD + 1];
  __shared__ float wred[8];
  __shared__ float pex[3][HD];
  pdl_launch_dependents();
  vb_trace_cta(16);   // CTA resident
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = n_head * HD;
  using T = bf16;
  T *kb = kcache + (int64_t)b * cache_seq_stride + (int64_t)h * cache_cap * HD;
  T *vb_ = vcache + (int64_t)b * cache_seq_stride + (int64_t)h * cache_cap * HD;
  const bool has_new = qp.part != nullptr;
  const int g = lane >> 3, j8 = (lane & 7) * 8;
  // The K rows of earlier tokens and the lengths do not depend on the kernels of THIS step that precede the
  // launch (the QKV projection only produces the current token), so the chunk geometry is worked out and the
  // first K batch is requested ahead of the dependency wait; the generated-token count is read again after
  // the wait and the batch re-requested should it have moved (it cannot when steps are separate graph launches).
  int kv_len, pos, c0, c1, n;
  uint4 kraw[U];
  auto setup = [&](int n_generated) {
    kv_len = max(1, min(text_len[b] + prompt_len[b] + n_generated, cache_cap));
    pos = kv_len - 1;  // cache row of the current token
    const int chunk = ((kv_len + nsplit - 1) / nsplit + 15) & ~15;
    c0 = sp * chunk;
    c1 = min(kv_len, c0 + chunk);
    n = max(0, c1 - c0);
    if (n > 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        kraw[u] = ldg_stream16(kb + (int64_t)(c0 + min(u * 16 + warp * 4 + g, n - 1)) * HD + j8);
    }
  };
  // (only with the fused QKV prologue: there the current token's row is served from shared memory; without it the
  // row was written to the cache by the kernel this launch depends on and nothing may be read ahead of the wait)
  const int n_gen_early = has_new ? n_gen[b] : -1;
  if (has_new) setup(n_gen_early);
  float qbias[3] = {0.f, 0.f, 0.f}, qc[3] = {0.f, 0.f, 0.f};
  if (tid < HD && has_new) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      qbias[j] = qp.bias[j * d + h * HD + tid];
      if (qp.fold.stats) qc[j] = qp.fold.c[j * d + h * HD + tid];
    }
  }
  pdl_wait();
  vb_trace(TR_ATTN * 2);
  vb_trace_cta(17);   // dependency resolved
  if (decode_row_finished(finished, b, h, sp, d, tid, nsplit, n_head, out, out16, part_o, part_ml)) return;
  int n_gen_now;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(n_gen_now) : "l"(n_gen + b) : "memory");
  if (n_gen_now != n_gen_early) setup(n_gen_now);  // uniform over the CTA
  if (has_new) {
    // q / k / v of the current token = sum of the projection's split-K partial tiles (+ folded LayerNorm, bias).  All 128
    // threads fetch: thread = (column c of the head, parity of the split), <= 3 splits x 3 columns each in ONE round
    // trip; the odd half hands its sums over through shared memory.  (A per-thread loop over the splits costs one
    // dependent L2 round trip per split.)
    const int c = tid & (HD - 1), hf = tid >> 6;
    float2 mraw = make_float2(0.f, 0.f);
    if (qp.fold.stats && hf == 0) mraw = ln_fold_moments_load(qp.fold, b, h);   // warps 0 and 1, complete
    float pv[3][3];   // (<= 6 splits in the one round trip; the chain uses 5)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float *p = qp.part + (int64_t)b * qp.ldp + j * d + h * HD + c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int s = hf + 2 * k;
        pv[j][k] = s < qp.splits ? __ldcg(p + (int64_t)s * 64 * qp.ldp) : 0.f;
      }
    }
    float mean = 0.f, rstd = 1.f;
    if (qp.fold.stats && hf == 0) ln_fold_moments_finish(qp.fold, mraw, mean, rstd);
    float acc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      acc[j] = (pv[j][0] + pv[j][1]) + pv[j][2];
      const float *p = qp.part + (int64_t)b * qp.ldp + j * d + h * HD + c;
      for (int s = 6 + hf; s < qp.splits; s += 2) acc[j] += __ldcg(p + (int64_t)s * 64 * qp.ldp);
      if (hf == 1) pex[j][c] = acc[j];
    }
    __syncthreads();
    if (hf == 0) {
      float a[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) a[j] = rstd * ((acc[j] + pex[j][c]) - mean * qc[j]) + qbias[j];
      qs[tid] = a[0] * 0.125f;
      const T k16 = from_f32<T>(a[1]), v16 = from_f32<T>(a[2]);
      knew[tid] = to_f32(k16);  // exactly what later steps will read back from the cache
      vnew[tid] = to_f32(v16);
      if (sp == 0) {
        kb[(int64_t)pos * HD + tid] = k16;
        vb_[(int64_t)pos * HD + tid] = v16;
      }
    }
  } else if (tid < HD) {
    qs[tid] = q[(int64_t)b * d + h * HD + tid] * 0.125f;
  }
  __syncthreads();

  // ---- scores: 8 lanes per key, 4 keys per warp-iteration, 16 keys per CTA-iteration ----
  float qf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qf[i] = qs[j8 + i];
  float lmax = -CUDART_INF_F;
  const bool new_here = has_new && pos >= c0 && pos < c1;  // the current token's key lives in smem
  // (refilling the registers of a half batch with the next batch's rows as soon as that half is consumed -- loads always
  //  in flight per warp -- measured the same: CTA duration 15.2 vs 15.3 us, profiles/round2_summary.md)
  for (int base = 0; base < n; base += 16 * U) {
    if (base > 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        kraw[u] = ldg_stream16(kb + (int64_t)(c0 + min(base + u * 16 + warp * 4 + g, n - 1)) * HD + j8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int key = base + u * 16 + warp * 4 + g;
      Vec16<bf16> v;
      v.raw = kraw[u];
      float kf[8];
      v.unpack(kf);
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(qf[i], kf[i], dot);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      if ((lane & 7) == 0 && key < n && !(new_here && c0 + key == pos)) {
        sc[key] = dot;
        lmax = fmaxf(lmax, dot);
      }
    }
  }
  // first batch of V rows in flight across the block softmax
  const int eg = (tid & 7) * 8, jl = tid >> 3;
  uint4 vraw[U];
  if (n > 0) {
#pragma unroll
    for (int u = 0; u < U; ++u) vraw[u] = ldg_stream16(vb_ + (int64_t)(c0 + min(u * 16 + jl, n - 1)) * HD + eg);
  }
  if (new_here && warp == 0) {  // score of the current token from the shared-memory key (never from the cache)
    float dot = 0.f;
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(qf[i], knew[j8 + i], dot);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    if (lane == 0) {
      sc[pos - c0] = dot;
      lmax = fmaxf(lmax, dot);
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) wred[warp] = lmax;
  __syncthreads();
  const float m = fmaxf(fmaxf(wred[0], wred[1]), fmaxf(wred[2], wred[3]));
  float lsum = 0.f;
  for (int i = tid; i < n; i += 128) {
    const float p = expf(sc[i] - m);
    sc[i] = p;
    lsum += p;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) wred[4 + warp] = lsum;
  __syncthreads();
  const float l = (wred[4] + wred[5]) + (wred[6] + wred[7]);

  // ---- O = P V : thread = (element group eg, key lane jl) --------------------------------
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int base = 0; base < n; base += 16 * U) {
    if (base > 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        vraw[u] = ldg_stream16(vb_ + (int64_t)(c0 + min(base + u * 16 + jl, n - 1)) * HD + eg);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int key = base + u * 16 + jl;
      const float pv = (key < n && !(new_here && c0 + key == pos)) ? sc[min(key, n - 1)] : 0.f;
      Vec16<bf16> v;
      v.raw = vraw[u];
      float vf[8];
      v.unpack(vf);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(pv, vf[i], acc[i]);
    }
  }
  if (new_here && jl == 0) {
    const float pn = sc[pos - c0];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pn, vnew[eg + i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[jl][eg + i] = acc[i];
  __syncthreads();
  if (tid < HD) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += red[r][tid];
    if (nsplit == 1) {
      out[(int64_t)b * d + h * HD + tid] = s / l;
      if (out16) out16[(int64_t)b * d + h * HD + tid] = __float2bfloat16_rn(s / l);
    } else {
      const int64_t pi = ((int64_t)b * n_head + h) * nsplit + sp;
      part_o[pi * HD + tid] = s;
      if (tid == 0) {
        part_ml[pi * 2] = n > 0 ? m : -CUDART_INF_F;
        part_ml[pi * 2 + 1] = n > 0 ? l : 0.f;
      }
    }
  }
  vb_trace(TR_ATTN * 2 + 1);
  vb_trace_cta(18);   // CTA done
}

__global__ void attn_decode_combine_kernel(const float *__restrict__ part_o,
                                           const float *__restrict__ part_ml, int n_head, int nsplit,
                                           float *__restrict__ out, bf16 *__restrict__ out16) {
  pdl_launch_dependents();
  pdl_wait();
  vb_trace(TR_COMBINE * 2);
  const int h = blockIdx.x, b = blockIdx.y, e = threadIdx.x;
  const int64_t p0 = ((int64_t)b * n_head + h) * nsplit;
  float m = -CUDART_INF_F;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, part_ml[(p0 + s) * 2]);
  float l = 0.f, o = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = part_ml[(p0 + s) * 2];
    if (ms == -CUDART_INF_F) continue;
    const float w = expf(ms - m);
    l += part_ml[(p0 + s) * 2 + 1] * w;
    o += part_o[(p0 + s) * HD + e] * w;
  }
  const float r = l > 0.f ? o / l : 0.f;  // l == 0: a finished utterance (every split empty)
  out[(int64_t)b * n_head * HD + h * HD + e] = r;
  if (out16) out16[(int64_t)b * n_head * HD + h * HD + e] = __float2bfloat16_rn(r);
}

static int decode_nsplit(int B, int n_head, int cache_cap) {
  const int forced = tune("VB_DECODE_NSPLIT", 0);
  if (forced > 0) return forced;
  int ns = (2 * sm_count() + B * n_head - 1) / (B * n_head);
  ns = max(1, min(ns, 32));
  ns = min(ns, max(1, cache_cap / 64));              // keep chunks >= 64 keys
  ns = max(ns, (cache_cap + kDecMaxChunk - 1) / kDecMaxChunk);  // chunk must fit the score buffer
  return ns;
}

size_t attn_decode_workspace(int B, int n_head, int head_dim, int cache_cap) {
  const int ns = decode_nsplit(B, n_head, cache_cap);
  return (size_t)B * n_head * ns * (head_dim + 2) * sizeof(float) + 256;
}

int launch_attn_decode(const float *q, const float *qkv_part, int qkv_splits, int qkv_ldp, const float *qkv_bias,
                       int B, int n_head, int head_dim, void *kcache, void *vcache, int dtype,
                       int64_t cache_seq_stride, int cache_cap, const int32_t *text_len, const int32_t *prompt_len,
                       const int32_t *n_gen, const int32_t *finished, float *out, void *out16, void *workspace,
                       bool pdl, cudaStream_t s, const LnFoldStats *fold) {
  VB_CHECK_ARG(head_dim == HD, "attn_decode: head_dim=%d, only 64 is built", head_dim);
  const int ns = decode_nsplit(B, n_head, cache_cap);
  VB_CHECK_ARG((cache_cap + ns - 1) / ns + 16 <= kDecMaxChunk, "attn_decode: cache_cap %d too large", cache_cap);
  float *part_o = (float *)workspace;
  float *part_ml = part_o + (size_t)B * n_head * ns * HD;
  QkvPartials qp{qkv_part, qkv_bias, qkv_splits, qkv_ldp, LnFoldStats{}};
  if (fold) qp.fold = *fold;
  dim3 grid(n_head, B, ns);
  if (dtype == VB_F32 || getenv("VB_ATTN_DECODE_1PASS") != nullptr) {  // fp32 parity path / single-pass variant
    if (dtype == VB_F32)
      VB_CUDA(launch_kernel(attn_decode_kernel<float>, grid, dim3(128), 0, s, pdl, q, qp, n_head, (float *)kcache,
                            (float *)vcache, cache_seq_stride, cache_cap, text_len, prompt_len, n_gen, finished, out,
                            (bf16 *)out16, part_o, part_ml, ns));
    else
      VB_CUDA(launch_kernel(attn_decode_kernel<bf16>, grid, dim3(128), 0, s, pdl, q, qp, n_head, (bf16 *)kcache,
                            (bf16 *)vcache, cache_seq_stride, cache_cap, text_len, prompt_len, n_gen, finished, out,
                            (bf16 *)out16, part_o, part_ml, ns));
  } else {
    // score buffer: the chunk of one split, rounded as the kernel rounds it (+16), in 1 KB steps
    const size_t sc_bytes = align_up((size_t)((cache_cap + ns - 1) / ns + 32) * sizeof(float), 1024);
    // VB_ATTN_CARVEOUT: shared-memory carve-out (percent) preferred for this kernel; -1 = the driver's choice.  72 % =
    // the 164 KB partition the projections of the chain run with: the driver's own pick for this (small) footprint
    // measured 1.5-3 % slower over the AR phase (the SMs re-partition on the way in and out of every attention launch),
    // the largest carve-out 25 % slower per launch (no L1 left for the loads in flight)
    static int carve_set[64];
    static bool carve_init = false;
    if (!carve_init) {
      for (int i = 0; i < 64; ++i) carve_set[i] = -2;
      carve_init = true;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    const int carve = tune("VB_ATTN_CARVEOUT", 72);
    if (carve_set[dev & 63] != carve) {
      const int want = carve >= 0 ? carve : (int)cudaSharedmemCarveoutDefault;
      if (carve >= 0 || carve_set[dev & 63] != -2)
        VB_CUDA(cudaFuncSetAttribute(attn_decode_2phase_pf_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, want));
      carve_set[dev & 63] = carve;
    }
    VB_CUDA(launch_kernel(attn_decode_2phase_pf_kernel<8>, grid, dim3(128), sc_bytes, s, pdl, q, qp, n_head,
                          (bf16 *)kcache, (bf16 *)vcache, cache_seq_stride, cache_cap, text_len, prompt_len, n_gen,
                          finished, out, (bf16 *)out16, part_o, part_ml, ns));
  }
  count_launch();
  if (ns > 1) {
    VB_CUDA(launch_kernel(attn_decode_combine_kernel, dim3(n_head, B), dim3(HD), 0, s, pdl, (const float *)part_o,
                          (const float *)part_ml, n_head, ns, out, (bf16 *)out16));
    count_launch();
  }
  return VB_OK;
}

}  // namespace vb

VB_API int vb_attention(const void *qkv, int dtype, int64_t M, int B, int n_head, int head_dim,
                        const int32_t *cu_seqlens, const int32_t *text_lens, const int32_t *seg1_lens, int seg1_start,
                        int max_seqlen, int mask_mode, void *out, void *kcache, void *vcache,
                        int64_t cache_seq_stride, int cache_cap, const uint8_t *dense_mask, int64_t dense_ld,
                        vb_stream_t stream) {
  return vb::launch_attention_varlen(qkv, dtype, M, B, n_head, head_dim, cu_seqlens, text_lens, seg1_lens, seg1_start,
                                     max_seqlen, mask_mode, out, kcache, vcache, cache_seq_stride, cache_cap,
                                     dense_mask, dense_ld, (cudaStream_t)stream);
}

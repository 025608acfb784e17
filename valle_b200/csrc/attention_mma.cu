// bf16 flash attention for packed ragged sequences (head_dim 64) on the warp-level tensor-core
// path (mma.sync m16n8k16, fp32 accumulate) -- the multi-query attention of the NAR passes, the AR
// prefill and the training forward in bf16 mode.  softmax(q k^T / 8 + mask) v with online softmax,
// K/V tiles double-buffered in shared memory by cp.async (16-byte, XOR-swizzled rows), Q kept in
// registers as A fragments, P re-used in registers as the A operand of P.V.
// Also fills the KV cache during the AR prefill.
//
// Reference arithmetic: F.multi_head_attention_forward (valle/modules/activation.py:408-427);
// masks: none for NAR (valle/models/valle.py:1125-1127), valle.py:1010-1033 for AR
// (kv_len(i) = max(S, i + 1)).
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vb {
namespace fa {

constexpr int HD = 64, BQ = 64, BKV = 64, kThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 64 x 64 bf16 tile, rows of 128 bytes = 8 chunks of 16 bytes; chunk index XOR-swizzled by (row & 7)
__device__ __forceinline__ int tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&p);
}

// load a 64-row x 64-col bf16 tile (rows row0.., column offset col0 of a [*, ld] matrix) with zero fill
__device__ __forceinline__ void load_tile(uint8_t *smem_tile, const bf16 *base, int64_t ld, int row0, int n_valid,
                                          int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * kThreads;  // 512 chunks
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (row0 + r) < n_valid;
    const bf16 *src = base + (int64_t)(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(smem_u32(smem_tile + tile_off(r, c)), src, ok);
  }
}

__global__ void __launch_bounds__(kThreads)
attn_varlen_mma_kernel(const bf16 *__restrict__ qkv, int n_head, const int32_t *__restrict__ cu_seqlens,
                       const int32_t *__restrict__ text_lens, const int32_t *__restrict__ seg1_lens,
                       int seg1_start, int mask_mode, bf16 *__restrict__ out,
                       bf16 *__restrict__ kcache, bf16 *__restrict__ vcache, int64_t cache_seq_stride,
                       int cache_cap, int tail_of_128) {
  __shared__ __align__(128) uint8_t sQ[BQ * 128];
  __shared__ __align__(128) uint8_t sK[2][BKV * 128];
  __shared__ __align__(128) uint8_t sV[2][BKV * 128];

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = cu_seqlens[b], L = cu_seqlens[b + 1] - r0;
  // tail_of_128: only the rows past the last full 128-row tile (the tcgen05 kernel covers the rest)
  const int q0 = (tail_of_128 ? (L & ~127) : 0) + blockIdx.x * BQ;
  if (q0 >= L) return;
  const int S = (mask_mode != VB_MASK_FULL) ? text_lens[b] : 0;
  const int c1 = (mask_mode >= VB_MASK_PADDED_AR) ? seg1_lens[b] : 0;
  const int d = n_head * HD;
  const int64_t ld = 3 * (int64_t)d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;

  const bf16 *qbase = qkv + (int64_t)r0 * ld + h * HD;
  const bf16 *kbase = qbase + d;
  const bf16 *vbase = qbase + 2 * d;
  const int q_hi = min(q0 + BQ, L);
  const int kv_max = (mask_mode == VB_MASK_VALLE_AR) ? max(S, q_hi) : L;
  const int n_tiles = (kv_max + BKV - 1) / BKV;

  load_tile(sQ, qbase, ld, q0, L, tid);
  load_tile(sK[0], kbase, ld, 0, L, tid);
  load_tile(sV[0], vbase, ld, 0, L, tid);
  cp_async_commit();

  // per-thread rows: g and g + 8 of this warp's 16-row slab
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  const RowMask lim_a = make_row_mask(mask_mode, row_a, L, S, seg1_start, c1);
  const RowMask lim_b = make_row_mask(mask_mode, row_b, L, S, seg1_start, c1);

  uint32_t qf[4][4];  // A fragments of Q: 4 k-steps over head_dim
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_a = -CUDART_INF_F, m_b = -CUDART_INF_F, l_a = 0.f, l_b = 0.f;
  const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

  for (int it = 0; it < n_tiles; ++it) {
    const int buf = it & 1;
    const int j0 = it * BKV;
    if (it + 1 < n_tiles) {  // prefetch next K/V tile
      load_tile(sK[buf ^ 1], kbase, ld, j0 + BKV, L, tid);
      load_tile(sV[buf ^ 1], vbase, ld, j0 + BKV, L, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 15), c = ks * 2 + (lane >> 4);
        ldmatrix_x4(smem_u32(sQ + tile_off(r, c)), qf[ks]);
      }
    }
    if (kcache != nullptr && j0 == q0) {  // this CTA owns cache rows [q0, q0+64)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * kThreads;
        const int r = idx >> 3, c = idx & 7;
        if (j0 + r < L) {
          const int64_t off = (int64_t)b * cache_seq_stride + ((int64_t)h * cache_cap + j0 + r) * HD + c * 8;
          *reinterpret_cast<uint4 *>(kcache + off) = *reinterpret_cast<const uint4 *>(sK[buf] + tile_off(r, c));
          *reinterpret_cast<uint4 *>(vcache + off) = *reinterpret_cast<const uint4 *>(sV[buf] + tile_off(r, c));
        }
      }
    }
    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
        uint32_t kf[4];
        // matrices: (keys np*16 + 0..7, hd ks*16 + 0..7), (same keys, hd +8), (keys +8, hd +0), (keys +8, hd +8)
        const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(smem_u32(sK[buf] + tile_off(r, c)), kf);
        mma_bf16(s[2 * np], qf[ks], kf[0], kf[1]);
        mma_bf16(s[2 * np + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // ---- mask + online softmax ----
    float mx_a = -CUDART_INF_F, mx_b = -CUDART_INF_F;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = j0 + nt * 8 + t4 * 2;
      s[nt][0] = lim_a.ok(c) ? s[nt][0] * sc : -CUDART_INF_F;
      s[nt][1] = lim_a.ok(c + 1) ? s[nt][1] * sc : -CUDART_INF_F;
      s[nt][2] = lim_b.ok(c) ? s[nt][2] * sc : -CUDART_INF_F;
      s[nt][3] = lim_b.ok(c + 1) ? s[nt][3] * sc : -CUDART_INF_F;
      mx_a = fmaxf(mx_a, fmaxf(s[nt][0], s[nt][1]));
      mx_b = fmaxf(mx_b, fmaxf(s[nt][2], s[nt][3]));
    }
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
    const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
    const float mu_a = mn_a == -CUDART_INF_F ? 0.f : mn_a, mu_b = mn_b == -CUDART_INF_F ? 0.f : mn_b;
    const float corr_a = exp2f(m_a - mu_a), corr_b = exp2f(m_b - mu_b);
    m_a = mn_a;
    m_b = mn_b;
    float rs_a = 0.f, rs_b = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mu_a);
      s[nt][1] = exp2f(s[nt][1] - mu_a);
      s[nt][2] = exp2f(s[nt][2] - mu_b);
      s[nt][3] = exp2f(s[nt][3] - mu_b);
      rs_a += s[nt][0] + s[nt][1];
      rs_b += s[nt][2] + s[nt][3];
    }
    l_a = l_a * corr_a + rs_a;  // per-thread partial row sums (quad-reduced at the end)
    l_b = l_b * corr_b + rs_b;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= corr_a;
      o[nt][1] *= corr_a;
      o[nt][2] *= corr_b;
      o[nt][3] *= corr_b;
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per k-step
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-wide head_dim n-tiles
        uint32_t vf[4];
        // transposed loads: (keys kk*16 + 0..7, hd np*16 + 0..7), (keys +8, hd +0), (keys +0, hd +8), (keys +8, hd +8)
        const int r = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int c = np * 2 + (lane >> 4);
        ldmatrix_x4_trans(smem_u32(sV[buf] + tile_off(r, c)), vf);
        mma_bf16(o[2 * np], pf, vf[0], vf[1]);
        mma_bf16(o[2 * np + 1], pf, vf[2], vf[3]);
      }
    }
    __syncthreads();  // tile fully consumed before the next prefetch overwrites the other buffer
  }
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
  const float inv_a = 1.f / l_a, inv_b = 1.f / l_b;
  bf16 *oa = out + (int64_t)(r0 + row_a) * d + h * HD + t4 * 2;
  bf16 *ob = out + (int64_t)(r0 + row_b) * d + h * HD + t4 * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (row_a < L) *reinterpret_cast<uint32_t *>(oa + nt * 8) = pack_bf16(o[nt][0] * inv_a, o[nt][1] * inv_a);
    if (row_b < L) *reinterpret_cast<uint32_t *>(ob + nt * 8) = pack_bf16(o[nt][2] * inv_b, o[nt][3] * inv_b);
  }
}

}  // namespace fa

int launch_attention_mma(const bf16 *qkv, int64_t M, int B, int n_head, const int32_t *cu_seqlens,
                         const int32_t *text_lens, const int32_t *seg1_lens, int seg1_start, int max_seqlen,
                         int mask_mode, bf16 *out, bf16 *kcache,
                         bf16 *vcache, int64_t cache_seq_stride, int cache_cap, int tail_of_128, cudaStream_t s) {
  if (M == 0 || B == 0) return VB_OK;
  dim3 grid(tail_of_128 ? 2 : (max_seqlen + fa::BQ - 1) / fa::BQ, n_head, B);
  fa::attn_varlen_mma_kernel<<<grid, fa::kThreads, 0, s>>>(qkv, n_head, cu_seqlens, text_lens, seg1_lens, seg1_start, mask_mode, out,
                                                          kcache, vcache, cache_seq_stride, cache_cap, tail_of_128);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

}  // namespace vb

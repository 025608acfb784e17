// Single-query decode attention over the bf16 KV cache with TMA-staged K/V tiles.
//
// The memory-bound kernel of the AR decode step (valle/models/valle.py:1012-1057 through
// F.multi_head_attention_forward, valle/modules/activation.py:408-427): one CTA per (head, utterance,
// kv-split).  The K rows [c0, c1) of one (b, h) are CONTIGUOUS in the cache layout [B, H, cap, 64], so a
// 64-key tile is ONE cp.async.bulk (8 KB, UBLKCP) into a shared-memory ring completed through an
// mbarrier; the ring first carries the K tiles, then the V tiles, so V is already in flight while the
// scores and the softmax of the chunk are computed.  Dots/softmax use warp shuffles; the (b, h) new
// key/value (split-K partial sums of the QKV projection, summed in fixed order in the prologue) never
// round-trips through global memory.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05_ptx.cuh"

namespace vb {
namespace adt {

using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int HD = 64, TILE = 64, NST = 8, kThreads = 128;
constexpr int kTileBytes = TILE * HD * 2;  // 8 KB
constexpr int kMaxChunk = 4096;

__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct QkvPartials {
  const float *part;
  const float *bias;
  int splits, ldp;
};

__global__ void __launch_bounds__(kThreads)
attn_decode_tma_kernel(const float *__restrict__ q, QkvPartials qp, int n_head, bf16 *__restrict__ kcache,
                       bf16 *__restrict__ vcache, int64_t cache_seq_stride, int cache_cap,
                       const int32_t *__restrict__ text_len, const int32_t *__restrict__ prompt_len,
                       const int32_t *__restrict__ n_gen, float *__restrict__ out, bf16 *__restrict__ out16,
                       float *__restrict__ part_o, float *__restrict__ part_ml, int nsplit, int sc_cap) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t *ring = smem_raw;                                            // [NST][8 KB]
  float *sc = reinterpret_cast<float *>(ring + NST * kTileBytes);      // [sc_cap]
  float *qs = sc + sc_cap;                                          // [64]
  float *knew = qs + HD;
  float *vnew = knew + HD;
  float *red = vnew + HD;                                              // [4][64]
  float *wred = red + 4 * HD;                                          // [8]
  uint64_t *full = reinterpret_cast<uint64_t *>(wred + 8);             // [NST]

  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = n_head * HD;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();
  int kv_len = text_len[b] + prompt_len[b] + n_gen[b];
  kv_len = max(1, min(kv_len, cache_cap));
  const int pos = kv_len - 1;
  const int chunk = ((kv_len + nsplit - 1) / nsplit + 15) & ~15;
  const int c0 = sp * chunk, c1 = min(kv_len, c0 + chunk);
  const int n = max(0, c1 - c0);
  const int nt = (n + TILE - 1) / TILE;   // tiles per matrix
  const int items = 2 * nt;               // K tiles then V tiles
  bf16 *kb = kcache + (int64_t)b * cache_seq_stride + (int64_t)h * cache_cap * HD;
  bf16 *vb_ = vcache + (int64_t)b * cache_seq_stride + (int64_t)h * cache_cap * HD;
  __syncthreads();  // barriers initialised

  auto issue = [&](int it) {  // thread 0 only
    const int t = it < nt ? it : it - nt;
    const int rows = min(TILE, n - t * TILE);
    const bf16 *src = (it < nt ? kb : vb_) + (int64_t)(c0 + t * TILE) * HD;
    uint64_t *bar = &full[it % NST];
    mbar_expect_tx(bar, rows * HD * 2);
    bulk_load(ring + (it % NST) * kTileBytes, src, rows * HD * 2, bar);
  };
  if (tid == 0)
    for (int it = 0; it < min(items, NST - 1); ++it) issue(it);

  // ---- prologue: q (and the new k/v) of the current token ---------------------------------------
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
        a[j] = acc + qp.bias[col];
      }
      qs[tid] = a[0] * 0.125f;
      const bf16 k16 = __float2bfloat16_rn(a[1]), v16 = __float2bfloat16_rn(a[2]);
      knew[tid] = __bfloat162float(k16);
      vnew[tid] = __bfloat162float(v16);
      if (sp == 0) {
        kb[(int64_t)pos * HD + tid] = k16;
        vb_[(int64_t)pos * HD + tid] = v16;
      }
    } else {
      qs[tid] = q[(int64_t)b * d + h * HD + tid] * 0.125f;
    }
  }
  __syncthreads();

  // ---- K phase: thread = (key = tid/2, half = tid&1), 32 dims each -------------------------------
  const int kkey = tid >> 1, khalf = tid & 1;
  float qf[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) qf[i] = qs[khalf * 32 + i];
  float lmax = -CUDART_INF_F;
  for (int t = 0; t < nt; ++t) {
    const int it = t;
    if (tid == 0 && it + NST - 1 < items) issue(it + NST - 1);
    mbar_wait(&full[it % NST], (it / NST) & 1);
    const uint8_t *tile = ring + (it % NST) * kTileBytes;
    const int key = t * TILE + kkey;
    float dot = 0.f;
    if (key < n) {
      if (has_new && c0 + key == pos) {
#pragma unroll
        for (int i = 0; i < 32; ++i) dot = fmaf(qf[i], knew[khalf * 32 + i], dot);
      } else {
        const uint4 *row = reinterpret_cast<const uint4 *>(tile + kkey * 128 + khalf * 64);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          Vec16<bf16> v;
          v.raw = row[c];
          float f[8];
          v.unpack(f);
#pragma unroll
          for (int i = 0; i < 8; ++i) dot = fmaf(qf[c * 8 + i], f[i], dot);
        }
      }
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    if (khalf == 0 && key < n) {
      sc[key] = dot;
      lmax = fmaxf(lmax, dot);
    }
    __syncthreads();  // tile consumed -> its ring slot may be refilled
  }
  lmax = warp_max(lmax);
  if (lane == 0) wred[warp] = lmax;
  __syncthreads();
  const float m = fmaxf(fmaxf(wred[0], wred[1]), fmaxf(wred[2], wred[3]));
  float lsum = 0.f;
  for (int i = tid; i < n; i += kThreads) {
    const float p = expf(sc[i] - m);
    sc[i] = p;
    lsum += p;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) wred[4 + warp] = lsum;
  __syncthreads();
  const float l = (wred[4] + wred[5]) + (wred[6] + wred[7]);

  // ---- V phase: thread = (dim pair = tid & 31, key quarter = tid >> 5) ---------------------------
  const int dp = tid & 31, kq = tid >> 5;
  float o0 = 0.f, o1 = 0.f;
  for (int t = 0; t < nt; ++t) {
    const int it = nt + t;
    if (tid == 0 && it + NST - 1 < items) issue(it + NST - 1);
    mbar_wait(&full[it % NST], (it / NST) & 1);
    const uint8_t *tile = ring + (it % NST) * kTileBytes;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int kbase = kq * 16 + g * 4;  // 4 keys at a time
      const int key0 = t * TILE + kbase;
      if (key0 >= n) break;
      const float4 p4 = *reinterpret_cast<const float4 *>(&sc[key0]);  // chunk is padded to 16 -> in bounds
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int key = key0 + u;
        if (key >= n) break;
        float v0, v1;
        if (has_new && c0 + key == pos) {
          v0 = vnew[dp * 2];
          v1 = vnew[dp * 2 + 1];
        } else {
          const uint32_t w = *reinterpret_cast<const uint32_t *>(tile + (kbase + u) * 128 + dp * 4);
          v0 = __uint_as_float(w << 16);
          v1 = __uint_as_float(w & 0xffff0000u);
        }
        o0 = fmaf(pv[u], v0, o0);
        o1 = fmaf(pv[u], v1, o1);
      }
    }
    __syncthreads();
  }
  red[kq * HD + dp * 2] = o0;
  red[kq * HD + dp * 2 + 1] = o1;
  __syncthreads();
  if (tid < HD) {
    const float s = (red[tid] + red[HD + tid]) + (red[2 * HD + tid] + red[3 * HD + tid]);
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
}

inline size_t smem_bytes(int sc_cap) {
  return NST * kTileBytes + (size_t)(sc_cap + 3 * HD + 4 * HD + 8) * sizeof(float) + NST * 8 + 64;
}

}  // namespace adt

int launch_attn_decode_tma(const float *q, const float *qkv_part, int qkv_splits, int qkv_ldp, const float *qkv_bias,
                           int B, int n_head, void *kcache, void *vcache, int64_t cache_seq_stride, int cache_cap,
                           const int32_t *text_len, const int32_t *prompt_len, const int32_t *n_gen, float *out,
                           void *out16, float *part_o, float *part_ml, int nsplit, bool pdl, cudaStream_t s) {
  const int sc_cap = (((cache_cap + nsplit - 1) / nsplit + 15) / 16 * 16 + 63) / 64 * 64;
  VB_CHECK_ARG(sc_cap <= adt::kMaxChunk, "attn_decode_tma: chunk %d too long", sc_cap);
  const size_t smem = adt::smem_bytes(sc_cap);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    VB_CUDA(cudaFuncSetAttribute(adt::attn_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  adt::QkvPartials qp{qkv_part, qkv_bias, qkv_splits, qkv_ldp};
  VB_CUDA(launch_kernel(adt::attn_decode_tma_kernel, dim3(n_head, B, nsplit), dim3(adt::kThreads), smem, s,
                        pdl, q, qp, n_head, (bf16 *)kcache, (bf16 *)vcache, cache_seq_stride, cache_cap, text_len,
                        prompt_len, n_gen, out, (bf16 *)out16, part_o, part_ml, nsplit, sc_cap));
  count_launch();
  return VB_OK;
}

}  // namespace vb

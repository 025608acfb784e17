// One AR decode step for 1..4 utterances as ONE persistent cooperative kernel (bf16 weights / KV cache, fp32
// activations): the small-batch / latency path of vb_ar_decode_step.
//
// With one to four rows a decode step is ~110 dependent launches of a few microseconds each (the split-K
// tensor-core chain of gemm_decode.cu pays for itself only when 16..64 rows share every weight byte).  Here one
// CTA per SM stays resident for the whole step and the stages are separated by grid barriers instead of kernel
// boundaries; a row of a weight matrix is owned by one warp (full K, no split-K, no partial sums), activations
// are fp32 vectors that every CTA re-reads from L2 after a barrier, LayerNorm and the split-KV combine are
// recomputed per CTA (a few KB) instead of being stages of their own:
//
//   per layer (valle/modules/transformer.py:297-334, activation.py:408-427):
//     S1  LN1(x) -> q,k,v = W_in . + b   (k, v appended to the cache at row pos, q to scratch)   | barrier
//     S2  single-query attention over the cache, (utterance, head, KV split) per CTA -> partial (m, l, o[64])
//                                                                                                | barrier
//     S3  combine the partials -> x += W_out . att + b                                           | barrier
//     S4  LN2(x) -> h = relu(W_1 . + b)                                                          | barrier
//     S5  x += W_2 . h + b                                                                       | barrier
//   then final LN -> logits = W_predict .  (valle.py:1039)                                       | barrier
//   then (greedy) argmax, stop rule, append, next embedding + PE (valle.py:1044-1057, 1013-1015), one CTA per row.
//
// Every stage ends by prefetching into L2 the weight rows the same warp owns in the next stage, so the HBM
// latency of the next stage overlaps the barrier.  61 barriers per step at d=1024/12L.
//
// STATUS: opt-in (VB_DECODE_PERSISTENT=1).  Measured on B200 at d=1024/12L: 0.388 ms/step at B=1 against 0.376
// for the PDL-chained launch sequence, 0.86 vs 0.39 at B=4 -- a stage costs ~6 us here (barrier + activation
// re-read + per-CTA LayerNorm + one latency-bound GEMV round), no less than a PDL-chained launch.  Kept as the
// tested skeleton (grid barrier, stage split, parity test) for a version with fewer, fatter stages.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vb {
namespace ps {

constexpr int kThreads = 512, kWarps = kThreads / 32;
constexpr int HD = 64;
constexpr int kMaxLayers = 16;
constexpr int kMaxChunk = 2048;  // keys per (utterance, head, split) work item

struct Params {
  vb_layer_params L[kMaxLayers];
  int n_layer, d, dff, H, B;
  const float *fn_w, *fn_b;
  // head / sampler
  const bf16 *predict_w;
  int n_vocab, eos_id, pe_rows, greedy, ld_logits;
  const float *audio_emb, *alpha, *pe;
  // loop state
  int tok_stride;
  const int32_t *text_len, *prompt_len, *max_new;
  int32_t *n_gen, *finished, *tokens;
  float *x, *logits;
  bf16 *kcache, *vcache;
  int64_t layer_stride, seq_stride;
  int cap;
  // scratch
  float *q, *hb, *part_o, *part_ml;
  int ns;
  unsigned *sync;  // [0] arrivals (monotonic), [1] value of [0] when the launch started
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_cg16(const void *p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
// all CTAs of the (cooperative) grid; `target` lives in thread 0 of every CTA
__device__ __forceinline__ void grid_barrier(unsigned *sync, unsigned &target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(sync, 1u);
    while ((int)(ld_acquire_u32(sync) - target) < 0) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void unpack8(const uint4 &raw, float (&f)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// L2 prefetch of the weight rows this warp owns in an upcoming GEMV stage
__device__ __forceinline__ void prefetch_rows(const bf16 *W, int n_rows, int K) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), GW = gridDim.x * kWarps;
  const int lines = (K * 2) >> 7;
  for (int row = gw; row < n_rows; row += GW) {
    const char *p = (const char *)(W + (int64_t)row * K);
    for (int i = lane; i < lines; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + ((int64_t)i << 7)));
  }
}

// out(row, b) = sum_k W[row, k] * xs[b][k]; one warp per row, rows interleaved over all warps of the grid
template <int NB, typename Epi>
__device__ __forceinline__ void gemv_rows(const bf16 *__restrict__ W, int n_rows, int K, const float *xs, int ldx,
                                          Epi epi) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), GW = gridDim.x * kWarps;
  for (int row = gw; row < n_rows; row += GW) {
    const bf16 *wr = W + (int64_t)row * K;
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll 4
    for (int c = lane * 8; c < K; c += 256) {
      float wf[8];
      unpack8(ldg_stream16(wr + c), wf);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 x0 = *reinterpret_cast<const float4 *>(xs + b * ldx + c);
        const float4 x1 = *reinterpret_cast<const float4 *>(xs + b * ldx + c + 4);
        acc[b] = fmaf(wf[0], x0.x, acc[b]);
        acc[b] = fmaf(wf[1], x0.y, acc[b]);
        acc[b] = fmaf(wf[2], x0.z, acc[b]);
        acc[b] = fmaf(wf[3], x0.w, acc[b]);
        acc[b] = fmaf(wf[4], x1.x, acc[b]);
        acc[b] = fmaf(wf[5], x1.y, acc[b]);
        acc[b] = fmaf(wf[6], x1.z, acc[b]);
        acc[b] = fmaf(wf[7], x1.w, acc[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) epi(row, b, acc[b]);
    }
  }
}

// block-wide sum of NB values (all threads get the result); red: [NB][kWarps] floats
template <int NB>
__device__ __forceinline__ void block_sum(float (&v)[NB], float *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[b] += __shfl_xor_sync(0xffffffffu, v[b], o);
    if (lane == 0) red[b * kWarps + warp] = v[b];
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[b * kWarps + w];
    v[b] = t;
  }
  __syncthreads();
}

// xs[b][:] = LayerNorm(x[b][:]) * gamma + beta for the B rows (two-pass moments, transformer.py:57-74)
template <int NB>
__device__ __forceinline__ void load_layernorm(const float *x, int B, int d, const float *gamma, const float *beta,
                                               float *xs, float *red) {
  const int tid = threadIdx.x;
  float s[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    s[b] = 0.f;
    if (b < B)
      for (int c = tid; c < d; c += kThreads) {
        const float v = __ldcg(x + (int64_t)b * d + c);
        xs[b * d + c] = v;
        s[b] += v;
      }
  }
  block_sum<NB>(s, red);
  float q[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    q[b] = 0.f;
    const float mean = s[b] / (float)d;
    if (b < B)
      for (int c = tid; c < d; c += kThreads) {
        const float t = xs[b * d + c] - mean;
        q[b] += t * t;
      }
  }
  block_sum<NB>(q, red);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float mean = s[b] / (float)d, rstd = rsqrtf(q[b] / (float)d + 1e-5f);
    for (int c = tid; c < d; c += kThreads)
      xs[b * d + c] = b < B ? (xs[b * d + c] - mean) * rstd * gamma[c] + beta[c] : 0.f;
  }
  __syncthreads();
}

template <int NB>
__global__ void __launch_bounds__(kThreads, 1) ar_step_persistent_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) float smem[];
  const int d = P.d, dff = P.dff, H = P.H, B = P.B;
  float *xs = smem;                         // [NB][max(d, dff)] activation vectors of the current stage
  float *sc = xs + (size_t)NB * max(d, dff);  // [kMaxChunk] attention scores
  float *red = sc + kMaxChunk;              // [64][HD + 1] reduction scratch (also block_sum)
  float *qs = red + 64 * (HD + 1);          // [HD]
  __shared__ int s_tok, s_pos;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned target = (tid == 0) ? P.sync[1] : 0u;

  for (int l = 0; l < P.n_layer; ++l) {
    const vb_layer_params &L = P.L[l];
    bf16 *kc = P.kcache + (int64_t)l * P.layer_stride;
    bf16 *vc = P.vcache + (int64_t)l * P.layer_stride;

    // ---- S1: LN1 -> QKV projection, KV append --------------------------------------------------------
    load_layernorm<NB>(P.x, B, d, L.norm1_w, L.norm1_b, xs, red);
    gemv_rows<NB>((const bf16 *)L.in_proj_w, 3 * d, d, xs, d, [&](int row, int b, float v) {
      if (b >= B) return;
      v += L.in_proj_b[row];
      const int part = row / d, c = row - part * d;
      if (part == 0) {
        P.q[(int64_t)b * d + c] = v;
      } else {
        const int h = c / HD, el = c - h * HD;
        int pos = P.text_len[b] + P.prompt_len[b] + P.n_gen[b] - 1;
        pos = max(0, min(pos, P.cap - 1));
        (part == 1 ? kc : vc)[(int64_t)b * P.seq_stride + ((int64_t)h * P.cap + pos) * HD + el] = __float2bfloat16_rn(v);
      }
    });
    prefetch_rows((const bf16 *)L.out_proj_w, d, d);
    grid_barrier(P.sync, target);

    // ---- S2: attention, one (utterance, head, KV split) per CTA -------------------------------------
    for (int item = blockIdx.x; item < B * H * P.ns; item += gridDim.x) {
      const int sp = item % P.ns, bh = item / P.ns, h = bh % H, b = bh / H;
      int kv_len = P.text_len[b] + P.prompt_len[b] + P.n_gen[b];
      kv_len = max(1, min(kv_len, P.cap));
      const int chunk = ((kv_len + P.ns - 1) / P.ns + 15) & ~15;
      const int c0 = sp * chunk, c1 = min(kv_len, c0 + chunk);
      const int n = max(0, c1 - c0);
      const bf16 *kb = kc + (int64_t)b * P.seq_stride + (int64_t)h * P.cap * HD;
      const bf16 *vb_ = vc + (int64_t)b * P.seq_stride + (int64_t)h * P.cap * HD;
      if (tid < HD) qs[tid] = __ldcg(P.q + (int64_t)b * d + h * HD + tid) * 0.125f;
      __syncthreads();
      // scores: 8 lanes per key, 4 keys per warp, 64 keys per CTA iteration
      const int g = lane >> 3, j8 = (lane & 7) * 8;
      float qf[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) qf[i] = qs[j8 + i];
      float lmax = -CUDART_INF_F;
      for (int base = 0; base < n; base += 4 * kWarps) {
        const int key = base + warp * 4 + g;
        float kf[8];
        unpack8(ld_cg16(kb + (int64_t)(c0 + min(key, n - 1)) * HD + j8), kf);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(qf[i], kf[i], dot);
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        if ((lane & 7) == 0 && key < n) {
          sc[key] = dot;
          lmax = fmaxf(lmax, dot);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
      if (lane == 0) red[warp] = lmax;
      __syncthreads();
      float m = red[0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) m = fmaxf(m, red[w]);
      __syncthreads();
      float lsum = 0.f;
      for (int i = tid; i < n; i += kThreads) {
        const float p = expf(sc[i] - m);
        sc[i] = p;
        lsum += p;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      if (lane == 0) red[warp] = lsum;
      __syncthreads();
      float lt = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) lt += red[w];
      __syncthreads();
      // O = P V : thread = (element group eg, key lane jl of 64)
      const int eg = (tid & 7) * 8, jl = tid >> 3;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int key = jl; key < n; key += 64) {
        float vf[8];
        unpack8(ld_cg16(vb_ + (int64_t)(c0 + key) * HD + eg), vf);
        const float p = sc[key];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) red[jl * (HD + 1) + eg + i] = acc[i];
      __syncthreads();
      if (tid < HD) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < 64; ++r) s += red[r * (HD + 1) + tid];
        P.part_o[(int64_t)item * HD + tid] = s;
        if (tid == 0) {
          P.part_ml[item * 2] = n > 0 ? m : -CUDART_INF_F;
          P.part_ml[item * 2 + 1] = n > 0 ? lt : 0.f;
        }
      }
      __syncthreads();
    }
    grid_barrier(P.sync, target);

    // ---- S3: combine the KV splits -> out-proj + residual ---------------------------------------------
    // split weights w_s = exp(m_s - m) / l per (utterance, head): one warp each, lanes = splits (ns <= 32)
    float *wsm = sc;  // [B * H][32]
    for (int bh = warp; bh < B * H; bh += kWarps) {
      const int i0 = bh * P.ns;
      const float ms = lane < P.ns ? __ldcg(P.part_ml + (i0 + lane) * 2) : -CUDART_INF_F;
      const float ls = lane < P.ns ? __ldcg(P.part_ml + (i0 + lane) * 2 + 1) : 0.f;
      float m = ms;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float w = ms == -CUDART_INF_F ? 0.f : expf(ms - m);
      float lt = ls * w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lt += __shfl_xor_sync(0xffffffffu, lt, o);
      wsm[bh * 32 + lane] = w / lt;
    }
    __syncthreads();
    for (int e = tid; e < NB * d; e += kThreads) {
      const int b = e / d, c = e - b * d, h = c / HD, el = c - h * HD;
      float o = 0.f;
      if (b < B) {
        const int bh = b * H + h;
        const float *po = P.part_o + (int64_t)bh * P.ns * HD + el;
#pragma unroll 4
        for (int s = 0; s < P.ns; ++s) o = fmaf(__ldcg(po + s * HD), wsm[bh * 32 + s], o);
      }
      xs[e] = o;
    }
    __syncthreads();
    gemv_rows<NB>((const bf16 *)L.out_proj_w, d, d, xs, d, [&](int row, int b, float v) {
      if (b >= B) return;
      float *xp = P.x + (int64_t)b * d + row;
      *xp = __ldcg(xp) + (v + L.out_proj_b[row]);
    });
    prefetch_rows((const bf16 *)L.lin1_w, dff, d);
    grid_barrier(P.sync, target);

    // ---- S4: LN2 -> FFN1 + ReLU ------------------------------------------------------------------------
    load_layernorm<NB>(P.x, B, d, L.norm2_w, L.norm2_b, xs, red);
    gemv_rows<NB>((const bf16 *)L.lin1_w, dff, d, xs, d, [&](int row, int b, float v) {
      if (b < B) P.hb[(int64_t)b * dff + row] = fmaxf(v + L.lin1_b[row], 0.f);
    });
    prefetch_rows((const bf16 *)L.lin2_w, d, dff);
    grid_barrier(P.sync, target);

    // ---- S5: FFN2 + residual ------------------------------------------------------------------------------
    for (int e = tid; e < NB * dff; e += kThreads) {
      const int b = e / dff;
      xs[e] = b < B ? __ldcg(P.hb + e) : 0.f;
    }
    __syncthreads();
    gemv_rows<NB>((const bf16 *)L.lin2_w, d, dff, xs, dff, [&](int row, int b, float v) {
      if (b >= B) return;
      float *xp = P.x + (int64_t)b * d + row;
      *xp = __ldcg(xp) + (v + L.lin2_b[row]);
    });
    if (l + 1 < P.n_layer)
      prefetch_rows((const bf16 *)P.L[l + 1].in_proj_w, 3 * d, d);
    else
      prefetch_rows(P.predict_w, P.n_vocab, d);
    grid_barrier(P.sync, target);
  }

  // ---- final LayerNorm -> ar_predict_layer ------------------------------------------------------------------
  load_layernorm<NB>(P.x, B, d, P.fn_w, P.fn_b, xs, red);
  gemv_rows<NB>(P.predict_w, P.n_vocab, d, xs, d, [&](int row, int b, float v) {
    if (b < B) P.logits[(int64_t)b * P.ld_logits + row] = v;
  });
  grid_barrier(P.sync, target);
  if (blockIdx.x == 0 && tid == 0) P.sync[1] = target;  // every CTA has arrived: [0] == target exactly

  // ---- greedy tail (valle.py:1044-1057, 1013-1015): one CTA per utterance -------------------------------------
  if (!P.greedy || blockIdx.x >= B) return;
  const int b = blockIdx.x;
  if (P.finished[b] != 0) return;
  float bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int i = tid; i < P.n_vocab; i += kThreads) {
    const float v = __ldcg(P.logits + (int64_t)b * P.ld_logits + i);
    if (v > bv || (v == bv && i < bi)) {
      bv = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  float *rv = red;
  int *ri = reinterpret_cast<int *>(red + kWarps);
  if (lane == 0) {
    rv[warp] = bv;
    ri[warp] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kWarps; ++w)
      if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) {
        bv = rv[w];
        bi = ri[w];
      }
    const int n_new = P.n_gen[b];
    const bool stop = (bi == P.eos_id) || (n_new > P.max_new[b]) || (n_new >= P.tok_stride);
    if (stop) {
      P.finished[b] = (n_new == 0) ? 2 : 1;
      s_tok = -1;
    } else {
      P.tokens[(int64_t)b * P.tok_stride + n_new] = bi;
      P.n_gen[b] = n_new + 1;
      s_tok = bi;
      s_pos = min(P.prompt_len[b] + n_new, P.pe_rows - 1);
    }
  }
  __syncthreads();
  const int tok = s_tok;
  if (tok < 0) return;
  const float a = P.alpha[0];
  const float *e = P.audio_emb + (int64_t)tok * d;
  const float *pp = P.pe + (int64_t)s_pos * d;
  for (int c = tid; c < d; c += kThreads) P.x[(int64_t)b * d + c] = __fadd_rn(e[c], __fmul_rn(a, pp[c]));
}

size_t smem_bytes(int NB, int d, int dff) {
  return ((size_t)NB * (size_t)std::max(d, dff) + kMaxChunk + 64 * (HD + 1) + HD) * sizeof(float);
}

}  // namespace ps

bool persistent_step_supported(const vb_decoder_desc &D, int B, int cache_cap) {
  if (D.wdtype != VB_BF16 || B < 1 || B > 4 || D.n_layer > ps::kMaxLayers) return false;
  if (D.d_model % 256 != 0 || D.d_ff % 256 != 0 || D.d_model / D.n_head != ps::HD) return false;
  const int ns = std::max(1, std::min(sm_count() / (B * D.n_head), 32));
  if (((cache_cap + ns - 1) / ns + 16) > ps::kMaxChunk) return false;
  if (ps::smem_bytes(4, D.d_model, D.d_ff) > 200 * 1024 || B * D.n_head * 32 > ps::kMaxChunk) return false;
  return getenv("VB_DECODE_PERSISTENT") != nullptr;
}

// scratch: q [B, d], hb [B, dff], partials [B*H*ns, HD + 2], 2 barrier words
size_t persistent_step_workspace(const vb_decoder_desc &D, int B) {
  const size_t items = (size_t)B * D.n_head * 32;
  return align_up((size_t)B * D.d_model * 4, 256) + align_up((size_t)B * D.d_ff * 4, 256) +
         align_up(items * (ps::HD + 2) * 4, 256) + 256;
}

int launch_persistent_step(const vb_decoder_desc &D, const vb_layer_params *layers, const vb_ar_head *head,
                           vb_ar_state *st, void *scratch, unsigned *sync, cudaStream_t s) {
  const int B = st->B, d = D.d_model, dff = D.d_ff;
  ps::Params P{};
  for (int l = 0; l < D.n_layer; ++l) P.L[l] = layers[l];
  P.n_layer = D.n_layer; P.d = d; P.dff = dff; P.H = D.n_head; P.B = B;
  P.fn_w = D.final_norm_w; P.fn_b = D.final_norm_b;
  P.predict_w = (const bf16 *)head->predict_w;
  P.n_vocab = head->n_vocab; P.eos_id = head->eos_id; P.pe_rows = head->pe_rows; P.greedy = head->greedy;
  P.ld_logits = (head->n_vocab + 3) & ~3;
  P.audio_emb = head->audio_emb; P.alpha = head->alpha; P.pe = head->pe;
  P.tok_stride = st->tok_stride;
  P.text_len = st->text_len; P.prompt_len = st->prompt_len; P.max_new = st->max_new;
  P.n_gen = st->n_gen; P.finished = st->finished; P.tokens = st->tokens;
  P.x = st->x_cur; P.logits = st->logits;
  P.kcache = (bf16 *)st->kcache; P.vcache = (bf16 *)st->vcache;
  P.layer_stride = st->cache_layer_stride; P.seq_stride = st->cache_seq_stride; P.cap = st->cache_cap;
  char *p = (char *)scratch;
  P.q = (float *)p; p += align_up((size_t)B * d * 4, 256);
  P.hb = (float *)p; p += align_up((size_t)B * dff * 4, 256);
  P.part_o = (float *)p;
  P.ns = std::max(1, std::min(sm_count() / (B * D.n_head), 32));
  P.part_ml = P.part_o + (size_t)B * D.n_head * P.ns * ps::HD;
  P.sync = sync;

  const int NB = B == 1 ? 1 : (B == 2 ? 2 : 4);
  auto kern = NB == 1 ? ps::ar_step_persistent_kernel<1>
                      : NB == 2 ? ps::ar_step_persistent_kernel<2> : ps::ar_step_persistent_kernel<4>;
  const size_t smem = ps::smem_bytes(NB, d, dff);
  static size_t attr_bytes[3] = {0, 0, 0};  // the attribute is a cap: only ever raise it (models of several sizes)
  const int ai = NB == 1 ? 0 : (NB == 2 ? 1 : 2);
  if (smem > attr_bytes[ai]) {
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes[ai] = smem;
  }
  int per_sm = 0;
  VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ps::kThreads, smem));
  VB_CHECK_ARG(per_sm >= 1, "persistent decode step: kernel does not fit an SM (smem %zu)", smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sm_count());
  cfg.blockDim = dim3(ps::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident or the launch fails (no silent deadlock)
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  VB_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  count_launch();
  return VB_OK;
}

}  // namespace vb

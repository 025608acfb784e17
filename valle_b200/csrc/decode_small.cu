// Small-batch (1..4 utterances) AR decoding as ONE persistent cooperative kernel that runs up to n_steps decode
// steps per launch: the latency path of VALLE.inference at batch 1 (BASELINE configs[1]; valle/models/valle.py:
// 1012-1057 loop, transformer.py:297-334 layer, activation.py:408-427 attention), bf16 weights / KV cache, fp32
// activations.
//
// With one row a decode step has ~336 MB to stream (51 us at the HBM roofline) but, as a chain of launches, costs
// ~110 dependent stages of ~3.5 us.  Here one CTA per SM stays resident; a stage boundary is a grid barrier, and
// what a stage needs from HBM is already in shared memory when the barrier drops:
//   * every CTA owns a fixed, contiguous block of output rows of each weight matrix; that block (<= 57 KB) is
//     fetched by ONE cp.async.bulk into a double-buffered shared-memory slot TWO stages ahead (weights do not
//     depend on activations), completion tracked by an mbarrier -- the weight stream never waits for a barrier;
//   * LayerNorm and the split-KV combine are recomputed per CTA from the few KB of activations in L2 instead of
//     being stages of their own; a matrix row is reduced by one warp over the full K (no split-K, no partials);
//   * steps are looped inside the kernel, so the weight pipeline runs across the step boundary and nothing is
//     re-launched per token; the host polls the stop flags once per launch.
//   per layer:  S1 LN1 + QKV (+KV append) | S2 attention, (row, head, KV split) per CTA | S3 combine + out-proj
//               + residual | S4 LN2 + FFN1 + ReLU | S5 FFN2 + residual;   then final LN + head | sampler.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace vb {
namespace sm {

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int HD = 64;
constexpr int kMaxLayers = 16;
constexpr int kMaxChunk = 2048;   // keys per (row, head, split) work item

struct Params {
  vb_layer_params L[kMaxLayers];
  int n_layer, d, dff, H, B;
  const float *fn_w, *fn_b;
  const bf16 *predict_w;
  int n_vocab, eos_id, pe_rows, ld_logits;
  const float *audio_emb, *alpha, *pe;
  int tok_stride;
  const int32_t *text_len, *prompt_len, *max_new;
  int32_t *n_gen, *finished, *tokens;
  float *x, *logits;
  bf16 *kcache, *vcache;
  int64_t layer_stride, seq_stride;
  int cap;
  float *q, *hb, *part_o, *part_ml;   // scratch: [B][d], [B][dff], [B*H*ns][64], [B*H*ns][2]
  int ns, n_steps, wbuf_bytes, barrier_mode;
  unsigned *sync;                      // grid-barrier counter, zeroed by the host before the launch
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#define grid_barrier(sync, target) grid_barrier_sync(sync, target, P.barrier_mode)

// the current token's K / V rows were written by other CTAs one barrier ago: read through L2
__device__ __forceinline__ uint4 ld_cg16(const void *p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void unpack8(const uint4 &raw, float (&f)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// weight stage `idx` of a step: 4 per layer (in-proj, out-proj, linear1, linear2) + the prediction head
struct WStage {
  const bf16 *W;
  int N, K;
};
__device__ __forceinline__ WStage wstage(const Params &P, int idx) {
  if (idx == 4 * P.n_layer) return WStage{P.predict_w, P.n_vocab, P.d};
  const vb_layer_params &L = P.L[idx >> 2];
  switch (idx & 3) {
    case 0: return WStage{(const bf16 *)L.in_proj_w, 3 * P.d, P.d};
    case 1: return WStage{(const bf16 *)L.out_proj_w, P.d, P.d};
    case 2: return WStage{(const bf16 *)L.lin1_w, P.dff, P.d};
    default: return WStage{(const bf16 *)L.lin2_w, P.d, P.dff};
  }
}
__device__ __forceinline__ int rows_per_cta(int N) { return (N + gridDim.x - 1) / gridDim.x; }

// out(row r of this CTA, b) = sum_k wsm[r][k] * xs[b][k]: one warp per row, lanes over k (16-byte weight reads)
// `pre(r)` fetches what the epilogue of row r needs from global memory (bias, residual) BEFORE the dot product, so
// the L2 round trip overlaps the reduction instead of following it
template <int NB, typename Pre, typename Epi>
__device__ __forceinline__ void gemv_smem(const bf16 *wsm, int rows, int K, const float *xs, Pre pre, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kMaxRows = 4;   // rows of one warp fetched up front (a CTA owns <= 32 rows of any matrix here)
  decltype(pre(0)) cs[kMaxRows];
#pragma unroll
  for (int i = 0; i < kMaxRows; ++i)
    if (warp + i * kWarps < rows) cs[i] = pre(warp + i * kWarps);
  int ri = 0;
  for (int r = warp; r < rows; r += kWarps, ++ri) {
    const bf16 *wr = wsm + (size_t)r * K;
    const auto c = ri < kMaxRows ? cs[ri < kMaxRows ? ri : 0] : pre(r);
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
    for (int c = lane * 8; c < K; c += 256) {
      float wf[8];
      unpack8(*reinterpret_cast<const uint4 *>(wr + c), wf);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 x0 = *reinterpret_cast<const float4 *>(xs + b * K + c);
        const float4 x1 = *reinterpret_cast<const float4 *>(xs + b * K + c + 4);
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
    for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) epi(r, b, acc[b], c);
    }
  }
}
template <int NB> struct RowConst {   // bias of the row + the residual-stream value of every batch row
  float bias;
  float res[NB];
};

// xs[b][:] = LayerNorm(x[b][:]) for the B rows (transformer.py:57-74); x read from L2 (written by other CTAs).
// The row stays in registers (d <= 256 * 8); sum and sum of squares in ONE block reduction (two barriers in all).
template <int NB>
__device__ __forceinline__ void load_layernorm(const float *x, int B, int d, const float *gamma, const float *beta,
                                               float *xs, float *red) {
  constexpr int MAXV = 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float v[NB][MAXV];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = tid + i * kThreads;
      v[b][i] = (b < B && c < d) ? __ldcg(x + (int64_t)b * d + c) : 0.f;
      s += v[b][i];
      q += v[b][i] * v[b][i];
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) {
      red[(2 * b) * kWarps + warp] = s;
      red[(2 * b + 1) * kWarps + warp] = q;
    }
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      s += red[(2 * b) * kWarps + w];
      q += red[(2 * b + 1) * kWarps + w];
    }
    const float mean = s / (float)d;
    const float rstd = rsqrtf(fmaxf(q / (float)d - mean * mean, 0.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = tid + i * kThreads;
      if (c < d) xs[b * d + c] = b < B ? (v[b][i] - mean) * rstd * gamma[c] + beta[c] : 0.f;
    }
  }
  __syncthreads();
}

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

template <int NB>
__global__ void __launch_bounds__(kThreads, 1) ar_steps_small_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int d = P.d, dff = P.dff, H = P.H, B = P.B;
  bf16 *wbuf[2] = {reinterpret_cast<bf16 *>(smem_raw), reinterpret_cast<bf16 *>(smem_raw + P.wbuf_bytes)};
  float *xs = reinterpret_cast<float *>(smem_raw + 2 * (size_t)P.wbuf_bytes);   // [NB][max(d, dff)]
  float *sc = xs + (size_t)NB * max(d, dff);                                     // [kMaxChunk] scores / [32][65] reduce
  float *red = sc + kMaxChunk + 64;                                              // [0,64) reductions, [64,128) query
  uint64_t *wbar = reinterpret_cast<uint64_t *>(red + 128);  // (red: 128 floats)                      // [2]
  __shared__ int s_all_done, s_tok, s_pos;
  __shared__ int s_kvpos[4], s_fin[4];   // per step: cache row of the current token, stop flag of every batch row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int stages_per_step = 4 * P.n_layer + 1;
  unsigned target = 0;
  int next_w = 0;   // next weight stage (global index) whose slot will be consumed

  if (tid == 0) {
    mbar_init(&wbar[0], 1);
    mbar_init(&wbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // issue the bulk copy of weight stage `g` (global index over all steps of this launch) into its slot
  auto issue = [&](int g) {
    if (tid != 0 || g >= stages_per_step * P.n_steps) return;
    const WStage w = wstage(P, g % stages_per_step);
    const int R = rows_per_cta(w.N), row0 = cta * R, rows = max(0, min(R, w.N - row0));
    const uint32_t bytes = (uint32_t)rows * (uint32_t)w.K * 2u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bytes > 0) {
      mbar_expect_tx(&wbar[g & 1], bytes);
      bulk_g2s(wbuf[g & 1], w.W + (size_t)row0 * w.K, bytes, &wbar[g & 1]);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&wbar[g & 1])) : "memory");
    }
  };
  // wait for weight stage g, run `body(wsm, rows, row0)`, then refill the slot with stage g + 2
  auto with_weights = [&](int g, auto body) {
    const WStage w = wstage(P, g % stages_per_step);
    const int R = rows_per_cta(w.N), row0 = cta * R, rows = max(0, min(R, w.N - row0));
    mbar_wait(&wbar[g & 1], (uint32_t)((g >> 1) & 1));
    body(wbuf[g & 1], rows, row0, w.K);
    __syncthreads();   // every warp is done with the slot
    issue(g + 2);
    next_w = g + 1;
  };
  issue(0);
  issue(1);

  for (int step = 0; step < P.n_steps; ++step) {
    const int g0 = step * stages_per_step;
    if (tid < 4) {
      int fin = 1, pos = 0;
      if (tid < B) {
        fin = __ldcg(P.finished + tid);
        pos = max(0, min(P.text_len[tid] + P.prompt_len[tid] + __ldcg(P.n_gen + tid) - 1, P.cap - 1));
      }
      s_fin[tid] = fin;
      s_kvpos[tid] = pos;
    }
    __syncthreads();
    if (tid == 0) s_all_done = (s_fin[0] != 0) && (s_fin[1] != 0) && (s_fin[2] != 0) && (s_fin[3] != 0);
    __syncthreads();
    if (s_all_done) break;   // uniform over the grid: `finished` only changes in the sampler, a barrier ago
    for (int l = 0; l < P.n_layer; ++l) {
      const vb_layer_params &LP = P.L[l];
      bf16 *kc = P.kcache + (int64_t)l * P.layer_stride;
      bf16 *vc = P.vcache + (int64_t)l * P.layer_stride;
      vb_trace(TR_LN * 2);
      // ---- S1: LN1 + in-proj; q to scratch, k / v appended to the cache (activation.py:408) ----
      load_layernorm<NB>(P.x, B, d, LP.norm1_w, LP.norm1_b, xs, red);
      if (l == 0) vb_trace(16);   // S1: LayerNorm done
      with_weights(g0 + 4 * l + 0, [&](const bf16 *wsm, int rows, int row0, int K) {
        if (l == 0) vb_trace(18);  // S1: weight slot ready
        gemv_smem<NB>(wsm, rows, K, xs, [&](int r) { return LP.in_proj_b[row0 + r]; },
                      [&](int r, int b, float v, float bias) {
          if (b >= B) return;
          const int n = row0 + r;
          v += bias;
          const int part = n / d, c = n - part * d;
          if (part == 0) {
            P.q[(int64_t)b * d + c] = v;
          } else if (s_fin[b] == 0) {
            const int h = c / HD, e = c - h * HD;
            (part == 1 ? kc : vc)[(int64_t)b * P.seq_stride + ((int64_t)h * P.cap + s_kvpos[b]) * HD + e] = __float2bfloat16_rn(v);
          }
        });
        if (l == 0) vb_trace(20);  // S1: rows done
      });
      if (l == 0) vb_trace(22);    // S1: next weights issued, about to arrive at the barrier
      grid_barrier(P.sync, target);
      // ---- S2: single-query attention, one (row, head, KV split) per CTA ----
      vb_trace(TR_ATTN * 2);
      {
        const int ns = P.ns, item = cta;
        if (item < B * H * ns) {
          const int sp = item % ns, bh = item / ns, h = bh % H, b = bh / H;
          const int kv_len = s_kvpos[b] + 1;
          const int chunk = ((kv_len + ns - 1) / ns + 15) & ~15;
          const int c0 = sp * chunk, n = max(0, min(kv_len, c0 + chunk) - c0);
          const bf16 *kb = kc + (int64_t)b * P.seq_stride + (int64_t)h * P.cap * HD;
          const bf16 *vb_ = vc + (int64_t)b * P.seq_stride + (int64_t)h * P.cap * HD;
          float *qs = red + 64;   // [64] staged query
          const int g8 = lane >> 3, j8 = (lane & 7) * 8;
          const int eg = (tid & 7) * 8, jl = tid >> 3;
          // the first 128 keys' K and V rows and the query in ONE L2 / HBM round trip (the cache rows do not depend on q)
          uint4 k_first[4], v_first[4];
          if (n > 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              k_first[u] = ld_cg16(kb + (int64_t)(c0 + min(u * 32 + warp * 4 + g8, n - 1)) * HD + j8);
              v_first[u] = ld_cg16(vb_ + (int64_t)(c0 + min(jl + u * 32, n - 1)) * HD + eg);
            }
          }
          if (tid < HD) qs[tid] = __ldcg(P.q + (int64_t)b * d + h * HD + tid) * 0.125f;
          __syncthreads();
          float qf[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) qf[i] = qs[j8 + i];
          float lmax = -CUDART_INF_F;
          for (int base0 = 0; base0 < n; base0 += 128) {
           uint4 kraw[4];
#pragma unroll
           for (int u = 0; u < 4; ++u)   // all loads of the batch in flight before the first use
             kraw[u] = base0 == 0 ? k_first[u]
                                  : ld_cg16(kb + (int64_t)(c0 + min(base0 + u * 32 + warp * 4 + g8, n - 1)) * HD + j8);
#pragma unroll
           for (int u = 0; u < 4; ++u) {
            const int key = base0 + u * 32 + warp * 4 + g8;
            float kf[8];
            unpack8(kraw[u], kf);
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
          }
          lmax = warp_max(lmax);
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
          lsum = warp_sum(lsum);
          if (lane == 0) red[warp] = lsum;
          __syncthreads();
          float lt = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) lt += red[w];
          // O = P V: thread = (8 head dims eg, key lane jl of 32)
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
          for (int key0 = jl; key0 < n; key0 += 128) {
            uint4 vraw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              vraw[u] = key0 == jl ? v_first[u] : ld_cg16(vb_ + (int64_t)(c0 + min(key0 + u * 32, n - 1)) * HD + eg);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int key = key0 + u * 32;
              float vf[8];
              unpack8(vraw[u], vf);
              const float pv = key < n ? sc[key] : 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] = fmaf(pv, vf[i], acc[i]);
            }
          }
          __syncthreads();                       // scores consumed: reuse `sc` as the [32][65] reduction tile
          float *rt = sc;
#pragma unroll
          for (int i = 0; i < 8; ++i) rt[jl * 65 + eg + i] = acc[i];
          __syncthreads();
          if (tid < HD) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) s += rt[r * 65 + tid];
            P.part_o[(int64_t)item * HD + tid] = s;
            if (tid == 0) {
              P.part_ml[(int64_t)item * 2] = n > 0 ? m : -CUDART_INF_F;
              P.part_ml[(int64_t)item * 2 + 1] = n > 0 ? lt : 0.f;
            }
          }
        }
      }
      grid_barrier(P.sync, target);
      // ---- S3: combine the KV splits (every CTA, from L2), out-proj + bias + residual ----
      vb_trace(TR_GEMM * 2);
      {
        // pass 1 (one L2 round trip): the (m, l) pair of every (row, head, split) -> normalised split weights
        float *wgt = sc;                                 // [B * H * ns]
        const int n_items = B * H * P.ns;
        for (int i = tid; i < n_items; i += kThreads) {
          const float2 ml = __ldcg(reinterpret_cast<const float2 *>(P.part_ml) + i);
          wgt[i] = ml.x;
          wgt[n_items + i] = ml.y;
        }
        __syncthreads();
        for (int bh = tid; bh < B * H; bh += kThreads) {
          float m = -CUDART_INF_F;
          for (int s = 0; s < P.ns; ++s) m = fmaxf(m, wgt[bh * P.ns + s]);
          float lt = 0.f;
          for (int s = 0; s < P.ns; ++s) {
            const float ms = wgt[bh * P.ns + s];
            const float w = (ms == -CUDART_INF_F) ? 0.f : __expf(ms - m);
            lt += wgt[n_items + bh * P.ns + s] * w;
            wgt[bh * P.ns + s] = w;
          }
          const float inv = lt > 0.f ? 1.f / lt : 0.f;
          for (int s = 0; s < P.ns; ++s) wgt[bh * P.ns + s] *= inv;
        }
        __syncthreads();
        // pass 2 (one round trip, independent loads): att[b][h*64+e] = sum_s w_s o_s[e]
        for (int i = tid; i < B * d; i += kThreads) {
          const int b = i / d, c = i - b * d, hh = c / HD, e = c - hh * HD;
          const int bh = b * H + hh;
          const float *po = P.part_o + (int64_t)bh * P.ns * HD + e;
          float o = 0.f;
#pragma unroll 4
          for (int s = 0; s < P.ns; ++s) o = fmaf(__ldcg(po + s * HD), wgt[bh * P.ns + s], o);
          xs[b * d + c] = o;
        }
        for (int i = B * d + tid; i < NB * d; i += kThreads) xs[i] = 0.f;
        __syncthreads();
      }
      with_weights(g0 + 4 * l + 1, [&](const bf16 *wsm, int rows, int row0, int K) {
        gemv_smem<NB>(wsm, rows, K, xs,
                      [&](int r) {
                        RowConst<NB> c;
                        c.bias = LP.out_proj_b[row0 + r];
#pragma unroll
                        for (int b = 0; b < NB; ++b) c.res[b] = b < B ? __ldcg(P.x + (int64_t)b * d + row0 + r) : 0.f;
                        return c;
                      },
                      [&](int r, int b, float v, const RowConst<NB> &c) {
                        if (b < B) P.x[(int64_t)b * d + row0 + r] = c.res[b] + v + c.bias;
                      });
      });
      grid_barrier(P.sync, target);
      // ---- S4: LN2 + linear1 + ReLU (transformer.py:332-334) ----
      vb_trace(TR_RELU * 2);
      load_layernorm<NB>(P.x, B, d, LP.norm2_w, LP.norm2_b, xs, red);
      with_weights(g0 + 4 * l + 2, [&](const bf16 *wsm, int rows, int row0, int K) {
        gemv_smem<NB>(wsm, rows, K, xs, [&](int r) { return LP.lin1_b[row0 + r]; },
                      [&](int r, int b, float v, float bias) {
                        if (b < B) P.hb[(int64_t)b * dff + row0 + r] = fmaxf(v + bias, 0.f);
                      });
      });
      grid_barrier(P.sync, target);
      // ---- S5: linear2 + bias + residual ----
      vb_trace(TR_COMBINE * 2);
      for (int i = tid; i < NB * dff; i += kThreads) xs[i] = (i < B * dff) ? __ldcg(P.hb + i) : 0.f;
      __syncthreads();
      with_weights(g0 + 4 * l + 3, [&](const bf16 *wsm, int rows, int row0, int K) {
        gemv_smem<NB>(wsm, rows, K, xs,
                      [&](int r) {
                        RowConst<NB> c;
                        c.bias = LP.lin2_b[row0 + r];
#pragma unroll
                        for (int b = 0; b < NB; ++b) c.res[b] = b < B ? __ldcg(P.x + (int64_t)b * d + row0 + r) : 0.f;
                        return c;
                      },
                      [&](int r, int b, float v, const RowConst<NB> &c) {
                        if (b < B) P.x[(int64_t)b * d + row0 + r] = c.res[b] + v + c.bias;
                      });
      });
      grid_barrier(P.sync, target);
    }
    // ---- final LayerNorm + ar_predict_layer (valle.py:1039) ----
    vb_trace(TR_FUSED * 2);
    load_layernorm<NB>(P.x, B, d, P.fn_w, P.fn_b, xs, red);
    with_weights(g0 + 4 * P.n_layer, [&](const bf16 *wsm, int rows, int row0, int K) {
      gemv_smem<NB>(wsm, rows, K, xs, [&](int) { return 0; },
                    [&](int r, int b, float v, int) {
                      if (b < B) P.logits[(int64_t)b * P.ld_logits + row0 + r] = v;
                    });
    });
    grid_barrier(P.sync, target);
    // ---- sampler: argmax, stop rule, append, next input row (valle.py:1044-1057, 1013-1015); CTA b per row ----
    vb_trace(TR_SAMPLE * 2);
    if (cta < B) {
      const int b = cta;
      float *xo = P.x + (int64_t)b * d;
      if (__ldcg(P.finished + b) != 0) {
        for (int c = tid; c < d; c += kThreads) xo[c] = 0.f;
      } else {
        ArgMax best{-CUDART_INF_F, 0x7fffffff};
        for (int i = tid; i < P.n_vocab; i += kThreads) best = better(best, ArgMax{__ldcg(P.logits + (int64_t)b * P.ld_logits + i), i});
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ArgMax t2;
          t2.v = __shfl_xor_sync(0xffffffffu, best.v, o);
          t2.i = __shfl_xor_sync(0xffffffffu, best.i, o);
          best = better(best, t2);
        }
        ArgMax *wb = reinterpret_cast<ArgMax *>(red);
        if (lane == 0) wb[warp] = best;
        __syncthreads();
        if (tid == 0) {
          ArgMax a = wb[0];
          for (int w = 1; w < kWarps; ++w) a = better(a, wb[w]);
          const int n_new = P.n_gen[b];
          const bool stop = (a.i == P.eos_id) || (n_new > P.max_new[b]) || (n_new >= P.tok_stride);
          if (stop) {
            P.finished[b] = (n_new == 0) ? 2 : 1;
            s_tok = -1;
          } else {
            P.tokens[(int64_t)b * P.tok_stride + n_new] = a.i;
            P.n_gen[b] = n_new + 1;
            s_tok = a.i;
            s_pos = min(P.prompt_len[b] + n_new, P.pe_rows - 1);
          }
        }
        __syncthreads();
        const int tok = s_tok;
        if (tok < 0) {
          for (int c = tid; c < d; c += kThreads) xo[c] = 0.f;
        } else {
          const float a = P.alpha[0];
          const float *e = P.audio_emb + (int64_t)tok * d, *p = P.pe + (int64_t)s_pos * d;
          for (int c = tid; c < d; c += kThreads) xo[c] = __fadd_rn(e[c], __fmul_rn(a, p[c]));
        }
      }
    }
    grid_barrier(P.sync, target);
  }
  // drain (early exit when every row has stopped): the copies already issued for the next two weight stages must land
  // before the CTA gives its shared memory back
  for (int g = next_w; g < min(next_w + 2, stages_per_step * P.n_steps); ++g) mbar_wait(&wbar[g & 1], (uint32_t)((g >> 1) & 1));
}

}  // namespace sm

bool decode_small_supported(const vb_decoder_desc &D, int B, int cache_cap) {
  if (D.wdtype != VB_BF16 || B < 1 || B > 4 || D.n_layer > sm::kMaxLayers || D.d_model % 256 != 0 || D.d_ff % 256 != 0 ||
      D.d_model > 2048)
    return false;
  if (tune("VB_DECODE_SMALL", 1) == 0) return false;
  const int G = sm_count();
  const int ns = std::max(1, G / (B * D.n_head));
  return (cache_cap + ns - 1) / ns + 16 <= sm::kMaxChunk;
}

static int small_wbuf_bytes(const vb_decoder_desc &D, int n_vocab, int G) {
  auto rows = [&](int N) { return (N + G - 1) / G; };
  size_t m = 0;
  m = std::max(m, (size_t)rows(3 * D.d_model) * D.d_model * 2);
  m = std::max(m, (size_t)rows(D.d_model) * D.d_model * 2);
  m = std::max(m, (size_t)rows(D.d_ff) * D.d_model * 2);
  m = std::max(m, (size_t)rows(D.d_model) * D.d_ff * 2);
  m = std::max(m, (size_t)rows(n_vocab) * D.d_model * 2);
  return (int)align_up(m, 128);
}

size_t decode_small_workspace(const vb_decoder_desc &D, int B) {
  const int G = sm_count();
  const int ns = std::max(1, G / (std::max(1, B) * D.n_head));
  return align_up((size_t)B * D.d_model * 4, 256) + align_up((size_t)B * D.d_ff * 4, 256) +
         align_up((size_t)B * D.n_head * ns * (sm::HD + 2) * 4, 256) + 512;
}

int launch_decode_small(const vb_decoder_desc &D, const vb_layer_params *layers, const vb_ar_head *head, vb_ar_state *st,
                        void *scratch, int n_steps, cudaStream_t s) {
  const int B = st->B, d = D.d_model, dff = D.d_ff, G = sm_count();
  VB_CHECK_ARG(decode_small_supported(D, B, st->cache_cap), "decode_small: unsupported configuration");
  VB_CHECK_ARG(head->greedy, "decode_small: greedy decoding only");
  sm::Params P{};
  for (int l = 0; l < D.n_layer; ++l) P.L[l] = layers[l];
  P.n_layer = D.n_layer; P.d = d; P.dff = dff; P.H = D.n_head; P.B = B;
  P.fn_w = D.final_norm_w; P.fn_b = D.final_norm_b;
  P.predict_w = (const bf16 *)head->predict_w;
  P.n_vocab = head->n_vocab; P.eos_id = head->eos_id; P.pe_rows = head->pe_rows;
  P.ld_logits = (head->n_vocab + 3) & ~3;
  P.audio_emb = head->audio_emb; P.alpha = head->alpha; P.pe = head->pe;
  P.tok_stride = st->tok_stride;
  P.text_len = st->text_len; P.prompt_len = st->prompt_len; P.max_new = st->max_new;
  P.n_gen = st->n_gen; P.finished = st->finished; P.tokens = st->tokens;
  P.x = st->x_cur; P.logits = st->logits;
  P.kcache = (bf16 *)st->kcache; P.vcache = (bf16 *)st->vcache;
  P.layer_stride = st->cache_layer_stride; P.seq_stride = st->cache_seq_stride; P.cap = st->cache_cap;
  P.ns = std::max(1, G / (B * D.n_head));
  P.n_steps = n_steps;
  P.barrier_mode = tune("VB_GRID_BARRIER", 2);
  P.wbuf_bytes = small_wbuf_bytes(D, head->n_vocab, G);
  char *p = (char *)scratch;
  P.q = (float *)p;        p += align_up((size_t)B * d * 4, 256);
  P.hb = (float *)p;       p += align_up((size_t)B * dff * 4, 256);
  P.part_o = (float *)p;   p += align_up((size_t)B * D.n_head * P.ns * sm::HD * 4, 256);
  P.part_ml = (float *)p;  p += align_up((size_t)B * D.n_head * P.ns * 2 * 4, 256);
  P.sync = (unsigned *)p;
  VB_CUDA(cudaMemsetAsync(P.sync, 0, 64 * sizeof(unsigned), s));
  const int NB = B == 1 ? 1 : (B == 2 ? 2 : 4);
  const size_t smem = 2 * (size_t)P.wbuf_bytes + (size_t)NB * std::max(d, dff) * 4 + (sm::kMaxChunk + 64) * 4 + 128 * 4 + 64;
  VB_CHECK_ARG(smem <= 220 * 1024, "decode_small: needs %zu bytes of shared memory", smem);
  const void *kern = NB == 1 ? (const void *)sm::ar_steps_small_kernel<1>
                             : NB == 2 ? (const void *)sm::ar_steps_small_kernel<2> : (const void *)sm::ar_steps_small_kernel<4>;
  VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void *args[] = {(void *)&P};
  VB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(sm::kThreads), args, smem, s));
  count_launch();
  return VB_OK;
}

}  // namespace vb

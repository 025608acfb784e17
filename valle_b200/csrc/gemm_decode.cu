// Skinny (decode) projections on the tensor cores: out[b, n] = epi( sum_k act[b,k] W[n,k] + bias[n] ),
// b < B <= 64 rows of the AR decode step, bf16 operands, fp32 accumulate.
//
// HBM-bound weight streaming (AI <= 64 FLOP/B): the weight matrix is the M=128-row operand of
// tcgen05.mma ("swap-AB"), the B <= 64 activation rows are the N=64 operand, so every weight byte is
// read exactly once per step at full TMA throughput and 148 SMs are filled by split-K:
//   grid = (N_out/128 tiles, S splits); CTA (t, s) streams W[t*128 .. +128, k-range(s)] through a
//   TMA/mbarrier ring into tcgen05.mma (M=128, N=64, K=16), accumulates in 64 TMEM columns, and
//   either applies the fused epilogue itself (S == 1: +bias, ReLU -> bf16, residual, QKV scatter) or
//   writes its fp32 partial tile to partials[split][b][n]; the CONSUMER kernel (residual + LayerNorm,
//   the KV-cache attention prologue, the sampler) sums the S partials in fixed order 0..S-1, which
//   keeps the result deterministic and costs no extra launch.
//   Programmatic dependent launch: barrier/TMEM setup and the first kStages WEIGHT tiles (which do
//   not depend on the previous kernel) are issued before griddepcontrol.wait, so weight streaming
//   overlaps the tail of the previous kernel in the CUDA graph.
//
// Replaces F.linear at valle/modules/activation.py:408 (in/out-proj), valle/modules/transformer.py:332-334
// (FFN) and valle/models/valle.py:1039 (ar_predict_layer) for the batched decode step.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05_ptx.cuh"

namespace vb {
namespace dg {

using namespace tc;

constexpr int TM = 128;      // weight rows per tile (UMMA M)
constexpr int TN = 64;       // activation rows (UMMA N)
constexpr int kStages = 6;   // (8 stages -- every weight tile of FFN1 / FFN2 in flight before the dependency wait --
                             // measured no faster: 5.6 vs 5.3-5.6 us per launch)
constexpr int kWBytes = TM * BK * 2;  // 16 KB
constexpr int kXBytes = TN * BK * 2;  // 8 KB
constexpr int kStageBytes = kWBytes + kXBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
constexpr int kThreads = 256;
constexpr int kTmemCols = 64;

struct Epi {
  int mode;  // DG_* below
  int N, B;  // valid output features / rows
  const float *bias;
  float *out_f32;      // [B, ld_out] (RESIDUAL: in/out; F32: out; QKV: q)
  bf16 *out_bf16;      // [B, ld_out] (RELU_BF16)
  int64_t ld_out;
  // QKV scatter
  int d, head_dim, cache_cap;
  bf16 *kcache, *vcache;
  int64_t cache_seq_stride;
  const int32_t *text_len, *prompt_len, *n_gen, *finished;
};

__device__ __forceinline__ void apply_epi(const Epi &e, int n, int b, float v) {
  if (e.bias) v += e.bias[n];
  if (e.mode == DG_F32) {
    e.out_f32[(int64_t)b * e.ld_out + n] = v;
  } else if (e.mode == DG_RESIDUAL) {
    float *o = e.out_f32 + (int64_t)b * e.ld_out + n;
    *o = *o + v;
  } else if (e.mode == DG_RELU_BF16) {
    e.out_bf16[(int64_t)b * e.ld_out + n] = __float2bfloat16_rn(fmaxf(v, 0.f));
  } else {
    const int part = n / e.d, c = n - part * e.d;
    if (part == 0) {
      e.out_f32[(int64_t)b * e.ld_out + c] = v;
    } else if (e.finished == nullptr || e.finished[b] == 0) {
      const int h = c / e.head_dim, el = c - h * e.head_dim;
      int pos = e.text_len[b] + e.prompt_len[b] + e.n_gen[b] - 1;
      pos = max(0, min(pos, e.cache_cap - 1));
      const int64_t off = (int64_t)b * e.cache_seq_stride + ((int64_t)h * e.cache_cap + pos) * e.head_dim + el;
      (part == 1 ? e.kcache : e.vcache)[off] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_decode_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                   int num_kb, float *__restrict__ partials, int ldp, Epi epi, KvPrefetch pf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + kStages * kStageBytes);
  uint64_t *full_bar = bars, *empty_bar = bars + kStages, *tmem_full = bars + 2 * kStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
  // k-block range of this split (balanced, contiguous)
  const int base = num_kb / splits, rem = num_kb % splits;
  const int kb0 = split * base + min(split, rem);
  const int nkb = base + (split < rem ? 1 : 0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // weights do not depend on the previous kernel: fill the ring with W tiles first ...
      const int pre = min(nkb, kStages);
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&full_bar[i], kStageBytes);
        tma_load_2d(&tmap_w, &full_bar[i], tiles + i * kStageBytes, (kb0 + i) * BK, tile * TM);
      }
      pdl_wait();  // ... the activations do
      vb_trace(TR_GEMM * 2);
      for (int i = 0; i < pre; ++i)
        tma_load_2d(&tmap_x, &full_bar[i], tiles + i * kStageBytes + kWBytes, (kb0 + i) * BK, 0);
      int stage = 0;
      uint32_t phase = 1;  // the ring has wrapped once
      for (int i = pre; i < nkb; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t *w_dst = tiles + stage * kStageBytes;
        mbar_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_2d(&tmap_w, &full_bar[stage], w_dst, (kb0 + i) * BK, tile * TM);
        tma_load_2d(&tmap_x, &full_bar[stage], w_dst + kWBytes, (kb0 + i) * BK, 0);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TM, TN);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t w_addr = smem_u32(tiles + stage * kStageBytes);
        const uint64_t adesc = make_smem_desc(w_addr);
        const uint64_t bdesc = make_smem_desc(w_addr + kWBytes);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) != 0);
        tcgen05_commit(&empty_bar[stage]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tcgen05_commit(tmem_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int nl = q * 32 + lane;  // feature within the tile
    const int n = tile * TM + nl;
    float v[TN];
    // idle until the accumulator is complete: pull a slice of an upcoming layer's KV cache into L2
    kv_prefetch(pf, (blockIdx.y * gridDim.x + blockIdx.x) * 4 + q, gridDim.x * gridDim.y * 4);
    pdl_wait();
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), r);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 32u, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[32 + i] = __uint_as_float(r[i]);
    } else {
#pragma unroll
      for (int i = 0; i < TN; ++i) v[i] = 0.f;
    }
    if (splits == 1) {
      // stage the tile through (now idle) pipeline shared memory so rows can be walked dynamically
      float *sv = reinterpret_cast<float *>(tiles);  // [TN][TM]
#pragma unroll
      for (int b = 0; b < TN; ++b) sv[b * TM + nl] = v[b];
      __syncwarp();
      if (n < epi.N)
        for (int b = 0; b < epi.B; ++b) apply_epi(epi, n, b, sv[b * TM + nl]);
    } else {
      // partials[split][b][n]: for a fixed row b consecutive lanes write consecutive features
      float *mine = partials + (int64_t)split * TN * ldp + n;
#pragma unroll
      for (int b = 0; b < TN; ++b) mine[(int64_t)b * ldp] = v[b];
    }
  }
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
  vb_trace(TR_GEMM * 2 + 1);
}


}  // namespace dg

size_t gemm_decode_workspace(int d_model, int d_ff) {
  // fp32 partials [splits][64][ldp], ldp = tiles * 128: the automatic split count keeps tiles * splits <= #SMs;
  // the forced / tuned counts of the decode chain (api.cu, VB_SPLITS_*) are capped at kMaxForcedSplits per
  // projection, whose widest output is max(3 * d_model, d_ff) features
  const size_t tiles_max = ((size_t)std::max(3 * d_model, d_ff) + dg::TM - 1) / dg::TM;
  const size_t slabs = std::max((size_t)sm_count() + 32, tiles_max * kMaxForcedSplits);
  return slabs * dg::TN * dg::TM * sizeof(float);
}

static int pick_splits(int tiles, int num_kb) {
  int s = sm_count() / tiles;
  s = max(1, min(s, num_kb / 2));
  return max(1, min(s, kMaxForcedSplits));   // the consumers keep up to kMaxForcedSplits slabs of a column in flight
}

int launch_gemm_decode(const bf16 *act, int B, int64_t ld_act, const bf16 *W, int N, int K, int force_splits,
                       const float *bias, int mode, float *out_f32, bf16 *out_bf16, int64_t ld_out,
                       const QkvScatter *qkv, float *partials, size_t partial_bytes, int *out_splits, int *out_ldp,
                       const KvPrefetch *pf, bool pdl, cudaStream_t s) {
  VB_CHECK_ARG(B >= 1 && B <= dg::TN, "gemm_decode: B=%d not in [1,64]", B);
  VB_CHECK_ARG(K % tc::BK == 0 && ld_act % 8 == 0, "gemm_decode: K %% 64 != 0 or unaligned activations");
  const int tiles = (N + dg::TM - 1) / dg::TM;
  const int num_kb = K / tc::BK;
  int splits = force_splits > 0 ? std::min(force_splits, std::min(kMaxForcedSplits, num_kb)) : pick_splits(tiles, num_kb);
  const int ldp = tiles * dg::TM;
  if (splits > 1)
    VB_CHECK_ARG(partials && partial_bytes >= (size_t)splits * dg::TN * ldp * sizeof(float),
                 "gemm_decode: partial buffer too small");
  if (out_splits) *out_splits = splits;
  if (out_ldp) *out_ldp = ldp;
  CUtensorMap tw, tx;
  VB_TRY(tc::make_tmap(&tw, W, N, K, K, dg::TM));
  VB_TRY(tc::make_tmap(&tx, act, B, K, ld_act, dg::TN));
  dg::Epi e{};
  e.mode = mode; e.N = N; e.B = B; e.bias = bias;
  e.out_f32 = out_f32; e.out_bf16 = out_bf16; e.ld_out = ld_out;
  if (mode == DG_QKV && splits == 1) {
    VB_CHECK_ARG(qkv != nullptr, "gemm_decode: qkv scatter parameters missing");
    e.d = qkv->d; e.head_dim = qkv->head_dim; e.cache_cap = qkv->cache_cap;
    e.kcache = (bf16 *)qkv->kcache; e.vcache = (bf16 *)qkv->vcache;
    e.cache_seq_stride = qkv->cache_seq_stride;
    e.text_len = qkv->text_len; e.prompt_len = qkv->prompt_len; e.n_gen = qkv->n_gen;
    e.finished = qkv->finished;
    e.out_f32 = qkv->q; e.ld_out = qkv->d;
  }
  static PerDeviceOnce once;
  if (once.first())
    VB_CUDA(cudaFuncSetAttribute(dg::gemm_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dg::kSmemBytes));
  KvPrefetch pf0{};
  if (pf) pf0 = *pf;
  VB_CUDA(launch_kernel(dg::gemm_decode_kernel, dim3(tiles, splits), dim3(dg::kThreads), dg::kSmemBytes, s, pdl, tw,
                        tx, num_kb, partials, ldp, e, pf0));
  count_launch();
  return VB_OK;
}

}  // namespace vb

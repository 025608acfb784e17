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
#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05_ptx.cuh"

namespace vb {
namespace dg {

using namespace tc;

constexpr int TM = 128;      // weight rows per tile (UMMA M)
constexpr int TN = 64;       // activation rows (UMMA N)
constexpr int kStagesMax = 6;  // pipeline depth is a template parameter (6, or 3 to leave room for a co-resident kernel)
constexpr int kWBytes = TM * BK * 2;  // 16 KB
constexpr int kXBytes = TN * BK * 2;  // 8 KB
constexpr int kStageBytes = kWBytes + kXBytes;
constexpr int smem_bytes(int stages) { return stages * kStageBytes + 1024 + 256; }
constexpr int kSmemBytes = smem_bytes(kStagesMax);
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
  const int32_t *text_len, *prompt_len, *n_gen;
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
    } else {
      const int h = c / e.head_dim, el = c - h * e.head_dim;
      int pos = e.text_len[b] + e.prompt_len[b] + e.n_gen[b] - 1;
      pos = max(0, min(pos, e.cache_cap - 1));
      const int64_t off = (int64_t)b * e.cache_seq_stride + ((int64_t)h * e.cache_cap + pos) * e.head_dim + el;
      (part == 1 ? e.kcache : e.vcache)[off] = __float2bfloat16_rn(v);
    }
  }
}

template <int kStages>
__global__ void __launch_bounds__(kThreads) __maxnreg__(kStages == 3 ? 96 : 208)
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
}


// ------------------------------------------------------------------------------------------------
// Cluster variant: the S split-K CTAs of one output tile form a thread-block cluster (1, S, 1).  Each
// CTA parks its fp32 partial tile in its own shared memory, the cluster synchronises once, and CTA r
// reduces batch rows [r*rows, (r+1)*rows) by reading the S partial tiles over distributed shared memory in
// fixed order 0..S-1 (deterministic), then applies the fused epilogue.  No partials in global memory, no
// consumer-side sums, no extra launch.
// ------------------------------------------------------------------------------------------------
constexpr int kStagesC = 8;
constexpr int kSmemBytesC = kStagesC * kStageBytes + 1024 + 256;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t raddr;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(cta_rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_decode_cluster_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                           int num_kb, Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + kStagesC * kStageBytes);
  uint64_t *full_bar = bars, *empty_bar = bars + kStagesC, *tmem_full = bars + 2 * kStagesC;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y, splits = gridDim.y;  // cluster = the `splits` CTAs of a tile
  const int base = num_kb / splits, rem = num_kb % splits;
  const int kb0 = split * base + min(split, rem);
  const int nkb = base + (split < rem ? 1 : 0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStagesC; ++i) {
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
  float *sp = reinterpret_cast<float *>(tiles);  // [TN rows b][TM features n] partial tile (after the MMAs)

  if (warp == 0) {
    if (lane == 0) {
      const int pre = min(nkb, kStagesC);
      for (int i = 0; i < pre; ++i) {  // weights first: they do not depend on the previous kernel
        mbar_expect_tx(&full_bar[i], kStageBytes);
        tma_load_2d(&tmap_w, &full_bar[i], tiles + i * kStageBytes, (kb0 + i) * BK, tile * TM);
      }
      pdl_wait();
      for (int i = 0; i < pre; ++i)
        tma_load_2d(&tmap_x, &full_bar[i], tiles + i * kStageBytes + kWBytes, (kb0 + i) * BK, 0);
      int stage = 0;
      uint32_t phase = 1;
      for (int i = pre; i < nkb; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t *w_dst = tiles + stage * kStageBytes;
        mbar_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_2d(&tmap_w, &full_bar[stage], w_dst, (kb0 + i) * BK, tile * TM);
        tma_load_2d(&tmap_x, &full_bar[stage], w_dst + kWBytes, (kb0 + i) * BK, 0);
        if (++stage == kStagesC) {
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
        if (++stage == kStagesC) {
          stage = 0;
          phase ^= 1;
        }
      }
      tcgen05_commit(tmem_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // partial tile: TMEM -> registers -> own shared memory, [b][n] so that lanes (n) are conflict free
    const int q = warp & 3;
    const int nl = q * 32 + lane;
    pdl_wait();
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tcgen05_fence_after();
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + hb * 32, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) sp[(hb * 32 + i) * TM + nl] = __uint_as_float(r[i]);
      }
    } else {
#pragma unroll 8
      for (int b = 0; b < TN; ++b) sp[b * TM + nl] = 0.f;
    }
  }
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (splits > 1) cluster_sync_all();  // every partial tile of this output tile is in shared memory
  if (warp >= 4) {
    const int nl = (warp & 3) * 32 + lane;
    const int n = tile * TM + nl;
    const int rows = (epi.B + splits - 1) / splits;
    const int b_lo = split * rows, b_hi = min(epi.B, b_lo + rows);
    const uint32_t sp_addr = smem_u32(sp);
    if (n < epi.N) {
      if (splits == 1) {
        for (int b = b_lo; b < b_hi; ++b) apply_epi(epi, n, b, sp[b * TM + nl]);
      } else {
        // remote shared-memory addresses of this feature column in every peer
        uint32_t peer[8];
#pragma unroll
        for (int s = 0; s < 8; ++s)
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                       : "=r"(peer[s])
                       : "r"(sp_addr + (uint32_t)nl * 4u), "r"((uint32_t)min(s, splits - 1)));
        for (int b = b_lo; b < b_hi; b += 2) {  // two rows x up to 8 peers = 16 independent DSMEM loads in flight
          float v[2][8];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t off = (uint32_t)(min(b + r, b_hi - 1) * TM) * 4u;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              v[r][s] = 0.f;
              if (s < splits) asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v[r][s]) : "r"(peer[s] + off) : "memory");
            }
          }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (b + r >= b_hi) break;
            float acc = v[r][0];
#pragma unroll
            for (int s = 1; s < 8; ++s)
              if (s < splits) acc += v[r][s];  // fixed order 0..S-1
            apply_epi(epi, n, b + r, acc);
          }
        }
      }
    }
  }
  if (splits > 1) cluster_sync_all();  // peers are done reading this CTA's shared memory
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

}  // namespace dg

size_t gemm_decode_workspace() {
  // fp32 partials [splits][64][ldp]: tiles * splits <= #SMs (or splits == 1), ldp = tiles * 128
  return (size_t)(sm_count() + 32) * dg::TN * dg::TM * sizeof(float);
}

static int pick_splits(int tiles, int num_kb) {
  int s = sm_count() / tiles;
  s = max(1, min(s, num_kb / 2));
  return max(1, min(s, 32));
}

int launch_gemm_decode(const bf16 *act, int B, int64_t ld_act, const bf16 *W, int N, int K, int force_splits,
                       const float *bias, int mode, float *out_f32, bf16 *out_bf16, int64_t ld_out,
                       const QkvScatter *qkv, float *partials, size_t partial_bytes, int *out_splits, int *out_ldp,
                       const KvPrefetch *pf, bool pdl, cudaStream_t s) {
  // opt-in, measured slower when applied to every projection (cluster co-scheduling defeats PDL overlap):
  // VB_DECODE_CLUSTER=1 -> all modes, VB_DECODE_CLUSTER_MODES=<bitmask of DG_* modes> -> selected projections
  static const int cluster_mask = getenv("VB_DECODE_CLUSTER") != nullptr
                                      ? 0xf
                                      : (getenv("VB_DECODE_CLUSTER_MODES") ? atoi(getenv("VB_DECODE_CLUSTER_MODES")) : 0);
  const bool cluster_reduce = ((cluster_mask >> mode) & 1) != 0;
  VB_CHECK_ARG(B >= 1 && B <= dg::TN, "gemm_decode: B=%d not in [1,64]", B);
  VB_CHECK_ARG(K % tc::BK == 0 && ld_act % 8 == 0, "gemm_decode: K %% 64 != 0 or unaligned activations");
  const int tiles = (N + dg::TM - 1) / dg::TM;
  const int num_kb = K / tc::BK;
  int splits = force_splits > 0 ? force_splits : pick_splits(tiles, num_kb);
  if (cluster_reduce) splits = min(splits, 8);  // portable cluster size
  const int ldp = tiles * dg::TM;
  if (splits > 1 && !cluster_reduce)
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
    e.out_f32 = qkv->q; e.ld_out = qkv->d;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VB_CUDA(cudaFuncSetAttribute(dg::gemm_decode_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dg::smem_bytes(6)));
    VB_CUDA(cudaFuncSetAttribute(dg::gemm_decode_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dg::smem_bytes(3)));
    VB_CUDA(cudaFuncSetAttribute(dg::gemm_decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dg::kSmemBytesC));
    attr_set = true;
  }
  if (cluster_reduce) {
    // the epilogue runs inside the kernel: QKV scatter parameters are needed for every split count
    if (mode == DG_QKV) {
      VB_CHECK_ARG(qkv != nullptr, "gemm_decode: qkv scatter parameters missing");
      e.d = qkv->d; e.head_dim = qkv->head_dim; e.cache_cap = qkv->cache_cap;
      e.kcache = (bf16 *)qkv->kcache; e.vcache = (bf16 *)qkv->vcache;
      e.cache_seq_stride = qkv->cache_seq_stride;
      e.text_len = qkv->text_len; e.prompt_len = qkv->prompt_len; e.n_gen = qkv->n_gen;
      e.out_f32 = qkv->q; e.ld_out = qkv->d;
    }
    VB_CUDA(launch_kernel_cluster(dg::gemm_decode_cluster_kernel, dim3(tiles, splits), dim3(dg::kThreads),
                                  dg::kSmemBytesC, s, pdl, dim3(1, splits, 1), tw, tx, num_kb, e));
    count_launch();
    if (out_splits) *out_splits = 1;  // nothing left for a consumer to sum
    return VB_OK;
  }
  KvPrefetch pf0{};
  if (pf) pf0 = *pf;
  // a 3-stage ring (72 KB) leaves shared memory for the KV-cache attention CTAs of a second micro-batch stream
  static const bool shallow = getenv("VB_DECODE_GEMM_STAGES") && atoi(getenv("VB_DECODE_GEMM_STAGES")) == 3;
  if (shallow)
    VB_CUDA(launch_kernel(dg::gemm_decode_kernel<3>, dim3(tiles, splits), dim3(dg::kThreads), dg::smem_bytes(3), s, pdl,
                          tw, tx, num_kb, partials, ldp, e, pf0));
  else
    VB_CUDA(launch_kernel(dg::gemm_decode_kernel<6>, dim3(tiles, splits), dim3(dg::kThreads), dg::smem_bytes(6), s, pdl,
                          tw, tx, num_kb, partials, ldp, e, pf0));
  count_launch();
  return VB_OK;
}

}  // namespace vb

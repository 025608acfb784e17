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
  int red;   // DG_RESIDUAL with split-K, no partials: 1 = every split adds its tile into out_f32 by a TMA bulk reduction
             // (order of the splits not fixed), 2 = the splits of a tile are a cluster and reduce over DSMEM in fixed order
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
                   const __grid_constant__ CUtensorMap tmap_red, int num_kb, float *__restrict__ partials, int ldp,
                   Epi epi, KvPrefetch pf) {
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
    if (epi.red) prefetch_tmap(&tmap_red);
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
    } else if (epi.red == 1) {
      // residual stream assembled in place: x[0:B, tile] += this split's tile (+ bias once, from split 0).  The tile
      // is staged row-major in the (now idle) pipeline shared memory and handed to the TMA engine as ONE bulk tensor
      // reduction (rows past B are clipped by the tensor map); per-lane red.global.add measured 2 us slower per launch.
      float *sv = reinterpret_cast<float *>(tiles);  // [TN][TM]
      const float bias = (split == 0 && epi.bias && n < epi.N) ? epi.bias[n] : 0.f;
#pragma unroll
      for (int b = 0; b < TN; ++b) sv[b * TM + nl] = v[b] + bias;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (q == 0 && lane == 0) {
        tma_reduce_add_2d(&tmap_red, sv, tile * TM, 0);
        tma_store_commit();
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // performed before this CTA retires (waiting only for
                                                                  // the shared-memory read measured the same)
      }
    } else if (epi.red == 2) {
      // deterministic variant: the CTAs of a tile (its `splits` splits) form a thread-block cluster; every CTA parks
      // its tile in shared memory, then CTA r adds up rows [r R, (r+1) R) of all the tiles over DSMEM in fixed order
      // 0..S-1 and updates the residual rows itself (plain read-modify-write, no atomics) -- after the barriers below
      float *sv = reinterpret_cast<float *>(tiles);  // [TN][TM]
#pragma unroll
      for (int b = 0; b < TN; ++b) sv[b * TM + nl] = v[b];
    } else {
      // partials[split][b][n]: for a fixed row b consecutive lanes write consecutive features
      float *mine = partials + (int64_t)split * TN * ldp + n;
#pragma unroll
      for (int b = 0; b < TN; ++b) mine[(int64_t)b * ldp] = v[b];
    }
  }
  __syncwarp();
  if (epi.red == 2 && splits > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp >= 4) {
      const int nl = (warp & 3) * 32 + lane, n = tile * TM + nl;
      const int R = (TN + splits - 1) / splits;
      const int b_lo = split * R, b_hi = min(min(b_lo + R, TN), epi.B);
      const uint32_t sv_addr = smem_u32(tiles) + (uint32_t)nl * 4u;
      const float bias = (epi.bias && n < epi.N) ? epi.bias[n] : 0.f;
      // remote address of this thread's column in every rank's tile (rank j == split j: the cluster is a y column)
      uint32_t ra[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra[j]) : "r"(sv_addr), "r"(min(j, splits - 1)));
#pragma unroll 2
      for (int b = b_lo; b < b_hi; ++b) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // all ranks' values requested before the first add (DSMEM latency ~200 clocks)
          t[j] = 0.f;
          if (j < splits)
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t[j]) : "r"(ra[j] + (uint32_t)(b * TM * 4)) : "memory");
        }
        float acc = t[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) acc += t[j];   // fixed order 0..S-1
        if (n < epi.N) {
          float *o = epi.out_f32 + (int64_t)b * epi.ld_out + n;
          *o = *o + (acc + bias);
        }
      }
    }
    // nobody leaves while its tile may still be read
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
  vb_trace(TR_GEMM * 2 + 1);
}

// ---- the same projection fed from the fp32 residual stream, LayerNorm folded into the weights ----------------
// LayerNorm(x) W^T = rstd (x (gamma o W)^T - mean c) + (beta W^T),  c[n] = sum_k gamma[k] W[n,k]: the tensor cores
// multiply the RAW rows by the pre-scaled weights, every CTA turns the fp32 rows of its k-range into the bf16
// 128B-swizzled operand tile itself (TMA brings the fp32 box, 8 warps convert it), and the moments of the rows
// (sum x, sum x^2 over the k-range) ride along as two more partial sums per split.  The consumer of the partials
// (attention prologue / ReLU reduce / sampler) applies rstd, mean, c and the folded bias -- the separate
// residual + LayerNorm launch between two projections (transformer.py:296-302,57-74) disappears from the chain.
constexpr int kStagesX = 4;   // the launcher picks the split count so that a CTA has <= 4 k-blocks where the split cap
                              // allows (d_model <= 4096): every tile in flight at once; beyond that the ring wraps
constexpr int kXfBytes = TN * BK * 4;  // 16 KB fp32 box
constexpr int kStageBytesX = kWBytes + kXBytes + kXfBytes;  // 40 KB
constexpr int kSmemBytesX = kStagesX * kStageBytesX + 1024 + 512;   // + alignment slack, barriers

constexpr int kThreadsX = 448;  // warp 0 TMA, 1 MMA, 2..9 converters (4..7 also the epilogue), 10..13 KV prefetch

// (register cap of two CTAs per SM: 72 registers, no spills -- with the 4-stage ring that leaves room for three CTAs of
//  the attention launch that follows to become resident, and fetch their first K rows, while this kernel still runs)
__global__ void __launch_bounds__(kThreadsX, 2)
gemm_decode_x_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_xf,
                     int num_kb, float *__restrict__ partials, int ldp, float *__restrict__ stats, KvPrefetch pf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + kStagesX * kStageBytesX);
  uint64_t *wfull = bars, *xfull = bars + kStagesX, *bfull = bars + 2 * kStagesX, *empty_bar = bars + 3 * kStagesX,
           *tmem_full = bars + 4 * kStagesX;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
  const int base = num_kb / splits, rem = num_kb % splits;
  const int kb0 = split * base + min(split, rem);
  const int nkb = base + (split < rem ? 1 : 0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_xf);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStagesX; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&xfull[i], 1);
      mbar_init(&bfull[i], 8);
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
      const int pre = min(nkb, kStagesX);
      for (int i = 0; i < pre; ++i) {  // weights: independent of the previous kernel
        mbar_expect_tx(&wfull[i], kWBytes);
        tma_load_2d(&tmap_w, &wfull[i], tiles + i * kStageBytesX, (kb0 + i) * BK, tile * TM);
      }
      pdl_wait();
      vb_trace(TR_GEMM * 2);
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&xfull[i], kXfBytes);
        tma_load_2d(&tmap_xf, &xfull[i], tiles + i * kStageBytesX + kWBytes + kXBytes, (kb0 + i) * BK, 0);
      }
      int stage = 0;
      uint32_t phase = 1;
      for (int i = pre; i < nkb; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t *dst = tiles + stage * kStageBytesX;
        mbar_expect_tx(&wfull[stage], kWBytes);
        tma_load_2d(&tmap_w, &wfull[stage], dst, (kb0 + i) * BK, tile * TM);
        mbar_expect_tx(&xfull[stage], kXfBytes);
        tma_load_2d(&tmap_xf, &xfull[stage], dst + kWBytes + kXBytes, (kb0 + i) * BK, 0);
        if (++stage == kStagesX) {
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
        mbar_wait(&wfull[stage], phase);
        mbar_wait(&bfull[stage], phase);
        tcgen05_fence_after();
        const uint32_t w_addr = smem_u32(tiles + stage * kStageBytesX);
        const uint64_t adesc = make_smem_desc(w_addr);
        const uint64_t bdesc = make_smem_desc(w_addr + kWBytes);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) != 0);
        tcgen05_commit(&empty_bar[stage]);
        if (++stage == kStagesX) {
          stage = 0;
          phase ^= 1;
        }
      }
      tcgen05_commit(tmem_full);
    }
    __syncwarp();
  } else if (warp < 10) {
    // ---- converters: warp cw owns rows cw*8 .. +8 of every k-block; lane = (row parity, float4 of the row) ----
    const int cw = warp - 2;
    const int rsub = lane >> 4, f = lane & 15;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    pdl_wait();
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&xfull[stage], phase);
      const uint8_t *xf = tiles + stage * kStageBytesX + kWBytes + kXBytes;
      uint8_t *xb = tiles + stage * kStageBytesX + kWBytes;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = cw * 8 + it * 2 + rsub;
        const float4 v = *reinterpret_cast<const float4 *>(xf + r * (BK * 4) + f * 16);
        s1[it] += (v.x + v.y) + (v.z + v.w);
        s2[it] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2[it]))));
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&p0);
        pk.y = *reinterpret_cast<uint32_t *>(&p1);
        // 128B swizzle of a K-major row: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
        *reinterpret_cast<uint2 *>(xb + r * 128 + ((((f >> 1) ^ (r & 7))) << 4) + (f & 1) * 8) = pk;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bfull[stage]);
      if (++stage == kStagesX) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (tile < kLnFoldMaxCopies && stats != nullptr) {  // moments of this split's k-range: stats[tile][split][row][2]
#pragma unroll
      for (int it = 0; it < 4; ++it) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          s1[it] += __shfl_xor_sync(0xffffffffu, s1[it], o);
          s2[it] += __shfl_xor_sync(0xffffffffu, s2[it], o);
        }
        if (f == 0) {
          const int r = cw * 8 + it * 2 + rsub;
          *reinterpret_cast<float2 *>(stats + (((int64_t)tile * splits + split) * TN + r) * 2) =
              make_float2(s1[it], s2[it]);
        }
      }
    }
    if (warp >= 4 && warp < 8) {  // ---- epilogue: fp32 partial tile of this split, 32 rows at a time ----
      const int q = warp & 3;
      const int n = tile * TM + q * 32 + lane;
      float *mine = partials + (int64_t)split * TN * ldp + n;
      if (nkb > 0) {
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        if (nkb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32), r);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) mine[(int64_t)(half * 32 + i) * ldp] = __uint_as_float(r[i]);
      }
    }
  } else {
    // idle warps: pull a slice of an upcoming layer's KV cache into L2 while the weight tiles stream
    kv_prefetch(pf, (blockIdx.y * gridDim.x + blockIdx.x) * 4 + (warp - 10), gridDim.x * gridDim.y * 4);
    pdl_wait();
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
                       const KvPrefetch *pf, bool pdl, cudaStream_t s, bool red_add) {
  VB_CHECK_ARG(B >= 1 && B <= dg::TN, "gemm_decode: B=%d not in [1,64]", B);
  VB_CHECK_ARG(!red_add || mode == DG_RESIDUAL, "gemm_decode: red_add needs the residual epilogue");
  VB_CHECK_ARG(K % tc::BK == 0 && ld_act % 8 == 0, "gemm_decode: K %% 64 != 0 or unaligned activations");
  const int tiles = (N + dg::TM - 1) / dg::TM;
  const int num_kb = K / tc::BK;
  int splits = force_splits > 0 ? std::min(force_splits, std::min(kMaxForcedSplits, num_kb)) : pick_splits(tiles, num_kb);
  const int ldp = tiles * dg::TM;
  if (splits > 1 && !red_add)
    VB_CHECK_ARG(partials && partial_bytes >= (size_t)splits * dg::TN * ldp * sizeof(float),
                 "gemm_decode: partial buffer too small");
  if (out_splits) *out_splits = red_add ? 1 : splits;  // nothing left for a consumer to add up
  if (out_ldp) *out_ldp = ldp;
  CUtensorMap tw, tx, tr;
  VB_TRY(tc::make_tmap(&tw, W, N, K, K, dg::TM));
  VB_TRY(tc::make_tmap(&tx, act, B, K, ld_act, dg::TN));
  if (red_add && splits > 1)
    VB_TRY(tc::make_tmap_f32_dense(&tr, out_f32, B, N, ld_out, dg::TN, dg::TM));
  else
    tr = tw;  // unused
  dg::Epi e{};
  e.mode = mode; e.N = N; e.B = B; e.bias = bias;
  e.red = (red_add && splits > 1) ? (tune("VB_RED_MODE", 1) == 2 && splits <= 8 ? 2 : 1) : 0;
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
  if (once.first()) {
    VB_CUDA(cudaFuncSetAttribute(dg::gemm_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dg::kSmemBytes));
  }
  KvPrefetch pf0{};
  if (pf) pf0 = *pf;
  VB_CUDA(launch_kernel_cluster(dg::gemm_decode_kernel, dim3(tiles, splits), dim3(dg::kThreads), dg::kSmemBytes, s, pdl,
                                dim3(1, e.red == 2 ? splits : 1, 1), tw, tx, tr, num_kb, partials, ldp, e, pf0));
  count_launch();
  return VB_OK;
}

// projection of the fp32 rows x[B, K] by LayerNorm-folded weights: fp32 partial tiles + the rows' moments per split
// (linear1 + ReLU finished inside this launch -- the splits of a tile as a thread-block cluster reducing over DSMEM --
//  was built and measured slower: 9.5 + 8.2 us for FFN1 + FFN2 against 5.0 + 2.7 + 7.4 us with relu_reduce_kernel)
int launch_gemm_decode_x(const float *x, int B, int64_t ldx, const bf16 *Wf, int N, int K, int force_splits,
                         float *partials, size_t partial_bytes, float *stats, int *out_splits, int *out_ldp,
                         int *out_copies, const KvPrefetch *pf, bool pdl, cudaStream_t s) {
  VB_CHECK_ARG(B >= 1 && B <= dg::TN, "gemm_decode_x: B=%d not in [1,64]", B);
  VB_CHECK_ARG(K % tc::BK == 0 && ldx % 4 == 0, "gemm_decode_x: K %% 64 != 0 or unaligned rows");
  const int tiles = (N + dg::TM - 1) / dg::TM;
  const int num_kb = K / tc::BK;
  int splits = force_splits > 0 ? std::min(force_splits, std::min(kMaxForcedSplits, num_kb)) : pick_splits(tiles, num_kb);
  // keep a CTA's k-range inside the ring where the split cap allows it (wider models: more splits rather than a
  // wrapping ring, whose later fp32 boxes would wait for the first MMAs)
  splits = std::max(splits, std::min(kMaxForcedSplits, (num_kb + dg::kStagesX - 1) / dg::kStagesX));
  const int ldp = tiles * dg::TM;
  VB_CHECK_ARG(partials && stats && partial_bytes >= (size_t)splits * dg::TN * ldp * sizeof(float),
               "gemm_decode_x: partial buffer too small");
  if (out_splits) *out_splits = splits;
  if (out_ldp) *out_ldp = ldp;
  if (out_copies) *out_copies = std::min(tiles, kLnFoldMaxCopies);
  CUtensorMap tw, tx;
  VB_TRY(tc::make_tmap(&tw, Wf, N, K, K, dg::TM));
  VB_TRY(tc::make_tmap_f32_dense(&tx, x, B, K, ldx, dg::TN, tc::BK));
  static PerDeviceOnce once;
  if (once.first())
    VB_CUDA(cudaFuncSetAttribute(dg::gemm_decode_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dg::kSmemBytesX));
  KvPrefetch pf0{};
  if (pf) pf0 = *pf;
  VB_CUDA(launch_kernel(dg::gemm_decode_x_kernel, dim3(tiles, splits), dim3(dg::kThreadsX), dg::kSmemBytesX, s, pdl,
                        tw, tx, num_kb, partials, ldp, stats, pf0));
  count_launch();
  return VB_OK;
}

// ---- LayerNorm folding (host API vb_ln_fold_build): wf[n,k] = bf16(W[n,k] gamma[k]), c[n] = sum_k wf[n,k],
//      dvec[n] = bias[n] + sum_k beta[k] W[n,k]; one warp per output feature --------------------------------------
__global__ void __launch_bounds__(256)
ln_fold_kernel(const bf16 *__restrict__ W, int N, int K, const float *__restrict__ gamma,
               const float *__restrict__ beta, const float *__restrict__ bias, bf16 *__restrict__ wf,
               float *__restrict__ c, float *__restrict__ dvec) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, ds = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = __bfloat162float(W[(int64_t)n * K + k]);
    const bf16 f = __float2bfloat16_rn(w * gamma[k]);
    wf[(int64_t)n * K + k] = f;
    cs += __bfloat162float(f);
    ds = fmaf(beta[k], w, ds);
  }
  cs = warp_sum(cs);
  ds = warp_sum(ds);
  if (lane == 0) {
    c[n] = cs;
    dvec[n] = ds + (bias ? bias[n] : 0.f);
  }
}
int launch_ln_fold(const bf16 *W, int N, int K, const float *gamma, const float *beta, const float *bias, bf16 *wf,
                   float *c, float *dvec, cudaStream_t s) {
  VB_CUDA(launch_kernel(ln_fold_kernel, dim3((N + 7) / 8), dim3(256), 0, s, false, W, N, K, gamma, beta, bias, wf, c,
                        dvec));
  count_launch();
  return VB_OK;
}

}  // namespace vb

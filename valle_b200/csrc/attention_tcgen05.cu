// bf16 flash attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), head_dim 64, packed
// ragged sequences -- the L x L attention of the 7 NAR passes and of the training forward
// (F.multi_head_attention_forward, valle/modules/activation.py:408-427; no mask for NAR
// valle/models/valle.py:1125-1127, key-padding / causal rules of valle.py:835-861,921-925).
//
// One CTA = 128 query rows of one (sequence, head); 320 threads:
//   warp 0    TMA producer: Q tile once, then K and V tiles (128 keys x 64, 128B-swizzle boxes of the
//             packed [M, 3d] qkv matrix) through a 2-stage mbarrier ring
//   warp 1    MMA issuer (one elected thread), tcgen05.mma with fp32 accumulators in TMEM; V is used as an
//             MN-major B operand exactly as TMA lands it, no transpose
//   warps 2-9 two softmax groups of 4 warps, one thread per query row (TMEM lane) -- see the pipeline
//             description above attn_tcgen05_pp_kernel
// two CTAs fit per SM (112 KB smem, 256 TMEM columns each) and interleave.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05_ptx.cuh"

namespace vb {
namespace fa5 {

using namespace tc;

constexpr int HD = 64, BQ = 128, BKV = 128;
constexpr int kThreads = 320;  // TMA warp, MMA warp, 2 x 4 softmax warps
constexpr int kQBytes = BQ * HD * 2;         // 16 KB
constexpr int kKBytes = BKV * HD * 2;        // 16 KB
constexpr int kStageBytes = 2 * kKBytes;     // K + V
constexpr int kStages = 2;
constexpr int kPBytes = BQ * BKV * 2;        // 32 KB (two 64-key K-blocks of [128 x 64])
constexpr int kSmemBytes = kQBytes + kStages * kStageBytes + kPBytes + 128 + 480;  // + barriers, TMEM slot
static_assert(2 * (kSmemBytes + 1024) <= 228 * 1024, "two CTAs per SM must fit");
constexpr int kTmemCols = 256;               // S_0, S_1, O_0, O_1: 64 columns each

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MN-major, 128-byte swizzle descriptor: a [K rows x 64 MN] tile stored as rows of 128 bytes
// (8 rows = one 1024-byte swizzle atom).  SBO = 1024 B between 8-row groups along K.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;  // leading byte offset (MN direction, unused for N = 64)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset (K direction)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor with B MN-major (bit 16)
__host__ __device__ constexpr uint32_t make_idesc_bmn(int M, int N) { return make_idesc(M, N) | (1u << 16); }

// chunk-local visibility of 32 consecutive keys starting at key index cb:
// valid(i) = i < a  or  b0 <= i < b1   (the two segments of RowMask clamped to the chunk)
struct ChunkMask {
  int a, b0, b1;
  __device__ __forceinline__ ChunkMask(const RowMask &rm, int cb) {
    a = min(max(rm.lim0 - cb, 0), 32);
    b0 = min(max(rm.s1 - cb, 0), 32);
    b1 = min(max(rm.hi1 - cb, 0), 32);
  }
  __device__ __forceinline__ bool any() const { return a > 0 || b1 > b0; }
  __device__ __forceinline__ bool has_seg1() const { return b1 > b0; }
  __device__ __forceinline__ bool ok(int i) const { return i < a || (i >= b0 && i < b1); }
};

// row maximum of this thread's 64 raw scores (thread = TMEM lane = query row, score buffer of its group)
template <bool kMask>
__device__ __forceinline__ float half_row_max(uint32_t taddr, const RowMask &rm, int jc0) {
  float mx = -CUDART_INF_F;
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    if (!kMask) {
      tmem_ld32(taddr + c0, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
    } else {
      const ChunkMask cm(rm, jc0 + c0);
      if (!__any_sync(0xffffffffu, cm.any())) continue;   // nobody in the warp sees these 32 keys
      tmem_ld32(taddr + c0, r);
      if (!__any_sync(0xffffffffu, cm.has_seg1())) {       // common case: one prefix [0, a)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < cm.a) mx = fmaxf(mx, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (cm.ok(i)) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
  }
  return mx;
}
// p = 2^(s*c - m) for 64 columns -> bf16 -> one 64-key k-block of the swizzled P tile; returns the sum
template <bool kMask>
__device__ __forceinline__ float half_row_p(uint32_t taddr, const RowMask &rm, int jc0, float sc, float m_use,
                                            uint8_t *blk_row, int row) {
  float rs = 0.f;
  const float nm = -m_use;
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    uint32_t pk[16];
    bool live = true;
    if (kMask) {
      const ChunkMask cm(rm, jc0 + c0);
      live = __any_sync(0xffffffffu, cm.any());
      if (live) {
        tmem_ld32(taddr + c0, r);
        if (!__any_sync(0xffffffffu, cm.has_seg1())) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= cm.a) r[i] = 0xff800000u;  // -inf -> p = 0
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!cm.ok(i)) r[i] = 0xff800000u;
        }
      }
    } else {
      tmem_ld32(taddr + c0, r);
    }
    if (live) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ex2(fmaf(__uint_as_float(r[i]), sc, nm));
        const float p1 = ex2(fmaf(__uint_as_float(r[i + 1]), sc, nm));
        rs += p0 + p1;
        __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<uint32_t *>(&pp);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int chunk = (c0 >> 3) + j;
      *reinterpret_cast<uint4 *>(blk_row + ((chunk ^ (row & 7)) << 4)) =
          make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
  }
  return rs;
}

// ------------------------------------------------------------------------------------------------
// Ping-pong pipeline: the 128-key tile is split into two 64-key halves and EACH HALF HAS ITS OWN
// softmax group (4 warps, one thread per query row), its own score buffer S_g, its own P_g buffer
// and its own output accumulator O_g in TMEM (columns: S_0 0..63, S_1 64..127, O_0 128..191,
// O_1 192..255).  Group g runs an independent online softmax over keys {128 t + 64 g ...}; the two
// partial results are merged once at the end (split-KV merge).  Nothing is exchanged per tile, the
// tensor pipe serves one group while the other is in its exp pass, and O stays in TMEM: P V
// accumulates (enable_input_d) and the accumulator is only rescaled when a row maximum has grown by
// more than 2^8 since the reference maximum was taken (exact: the reference point cancels in O / l).
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
attn_tcgen05_pp_kernel(const __grid_constant__ CUtensorMap tmap, int n_head, const int32_t *__restrict__ cu_seqlens,
                       const int32_t *__restrict__ text_lens, const int32_t *__restrict__ seg1_lens, int seg1_start,
                       int mask_mode, bf16 *__restrict__ out, int skip_partial) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t *sQ = smem;
  uint8_t *sKV = sQ + kQBytes;                       // [stage][K | V]
  uint8_t *sP = sKV + kStages * kStageBytes;         // [group][128 rows x 128 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sP + kPBytes);
  uint64_t *q_full = bars;              // 1
  uint64_t *kv_full = bars + 1;         // [kStages]
  uint64_t *kv_empty = kv_full + kStages;
  uint64_t *s_full = kv_empty + kStages;   // [2] S_g ready in TMEM
  uint64_t *p_full = s_full + 2;           // [2] P_g in smem, S_g consumed (4 warp arrivals)
  uint64_t *o_done = p_full + 2;           // [2] P_g V accumulated: P_g buffer free, O_g stable
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_done + 2);
  static_assert((1 + 2 * kStages + 6) * 8 + 8 <= 96, "barrier block");

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = cu_seqlens[b], L = cu_seqlens[b + 1] - r0;
  int q0 = blockIdx.x * BQ;
  if (q0 >= L) return;
  if (q0 + BQ > L) {
    if (skip_partial == 1) return;
    if (skip_partial == 2 && L >= BQ) q0 = L - BQ;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = n_head * HD;
  const int S = (mask_mode != VB_MASK_FULL) ? text_lens[b] : 0;
  const int c1 = (mask_mode >= VB_MASK_PADDED_AR) ? seg1_lens[b] : 0;
  const int q_hi = min(q0 + BQ, L);
  const int kv_max = (mask_mode == VB_MASK_VALLE_AR) ? max(S, q_hi) : L;
  const int n_tiles = (kv_max + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&o_done[g], 1);
    }
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
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_2d(&tmap, q_full, sQ, h * HD, r0 + q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t *dst = sKV + stage * kStageBytes;
        mbar_expect_tx(&kv_full[stage], kStageBytes);
        tma_load_2d(&tmap, &kv_full[stage], dst, d + h * HD, r0 + t * BKV);
        tma_load_2d(&tmap, &kv_full[stage], dst + kKBytes, 2 * d + h * HD, r0 + t * BKV);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(BQ, 64);       // S_g = Q K_g^T : both K-major, N = 64 keys
      constexpr uint32_t idesc_o = make_idesc_bmn(BQ, HD);   // O_g += P_g V_g : A K-major, B MN-major
      constexpr int kHalfBytes = 64 * HD * 2;                // 64 keys of a K or V tile = 8 swizzle atoms
      const uint64_t qdesc = make_smem_desc(smem_u32(sQ));
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tcgen05_fence_after();
      for (int g = 0; g < 2; ++g) {
        const uint64_t kdesc = make_smem_desc(smem_u32(sKV + g * kHalfBytes));
#pragma unroll
        for (int k = 0; k < HD / UMMA_K; ++k)
          umma_bf16(tmem_base + g * 64, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0);
        tcgen05_commit(&s_full[g]);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int vstage = stage;
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_full[g], t & 1);             // P_g(t) written, S_g(t) consumed
          tcgen05_fence_after();
          if (t + 1 < n_tiles) {                    // S_g(t+1) first: its softmax group waits on it
            if (g == 0) {
              mbar_wait(&kv_full[stage], phase);
              tcgen05_fence_after();
            }
            const uint64_t kdesc = make_smem_desc(smem_u32(sKV + stage * kStageBytes + g * kHalfBytes));
#pragma unroll
            for (int k = 0; k < HD / UMMA_K; ++k)
              umma_bf16(tmem_base + g * 64, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0);
            tcgen05_commit(&s_full[g]);
          }
          const uint64_t pdesc = make_smem_desc(smem_u32(sP + g * kQBytes));
          const uint64_t vdesc = make_smem_desc_mn(smem_u32(sKV + vstage * kStageBytes + kKBytes + g * kHalfBytes));
#pragma unroll
          for (int k = 0; k < 64 / UMMA_K; ++k)
            umma_bf16(tmem_base + 128 + g * 64, pdesc + (uint64_t)(k * 2), vdesc + (uint64_t)(k * (2048 >> 4)),
                      idesc_o, (t > 0 || k > 0) ? 1u : 0u);
          tcgen05_commit(&o_done[g]);
        }
        tcgen05_commit(&kv_empty[vstage]);  // K(t), V(t) free once these MMAs retire
      }
    }
    __syncwarp();
  } else {
    // ===== softmax group g: one thread per query row, keys 128 t + 64 g .. + 63 =====
    const int quarter = warp & 3;                  // TMEM lanes this warp may touch
    const int g = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int qr = q0 + row;
    const RowMask rm = make_row_mask(mask_mode, qr, L, S, seg1_start, c1);
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + g * 64;
    const uint32_t o_addr = tmem_base + lane_off + 128 + g * 64;
    uint8_t *p_row = sP + g * kQBytes + row * 128;
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m_ref = -CUDART_INF_F, l = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int jc0 = t * BKV + g * 64;
      if (lane == 0) mbar_wait(&s_full[g], t & 1);
      __syncwarp();
      tcgen05_fence_after();
      const bool interior = __all_sync(0xffffffffu, jc0 + 64 <= rm.lim0);
      const float mx = interior ? half_row_max<false>(s_addr, rm, jc0) : half_row_max<true>(s_addr, rm, jc0);
      const float m_cand = fmaxf(m_ref, mx * sc);
      const bool need = (m_cand - m_ref) > 8.f;     // NaN (-inf - -inf) compares false
      if (t > 0) {                                  // P_g buffer free, O_g stable
        if (lane == 0) mbar_wait(&o_done[g], (t - 1) & 1);
        __syncwarp();
        tcgen05_fence_after();
      }
      if (__any_sync(0xffffffffu, need)) {
        const float corr = (m_cand == m_ref) ? 1.f : ex2(m_ref - m_cand);
        if (t > 0) {
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(o_addr + c0, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            tmem_st32(o_addr + c0, r);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        l *= corr;
        m_ref = m_cand;
      }
      const float m_use = m_ref == -CUDART_INF_F ? 0.f : m_ref;
      l += interior ? half_row_p<false>(s_addr, rm, jc0, sc, m_use, p_row, row)
                    : half_row_p<true>(s_addr, rm, jc0, sc, m_use, p_row, row);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P visible to the tensor-core proxy
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }
    // ---- merge the two key halves: out = (O_0 w_0 + O_1 w_1) / (l_0 w_0 + l_1 w_1), w_g = 2^(m_g - max) ----
    if (lane == 0) {
      mbar_wait(&o_done[0], (n_tiles - 1) & 1);
      mbar_wait(&o_done[1], (n_tiles - 1) & 1);     // in-order retirement: every MMA of this CTA is done,
    }                                               // so the Q tile can carry the (m, l) exchange
    __syncwarp();
    tcgen05_fence_after();
    float2 *xch = reinterpret_cast<float2 *>(sQ);
    xch[g * BQ + row] = make_float2(m_ref, l);
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
    const float2 e0 = xch[row], e1 = xch[BQ + row];
    const float mm = fmaxf(e0.x, e1.x);
    const float w0 = e0.x == -CUDART_INF_F ? 0.f : ex2(e0.x - mm);
    const float w1 = e1.x == -CUDART_INF_F ? 0.f : ex2(e1.x - mm);
    const float inv = 1.f / (e0.y * w0 + e1.y * w1);
    const float f0 = w0 * inv, f1 = w1 * inv;
    uint32_t a[32], c[32];
    tmem_ld32(tmem_base + lane_off + 128 + g * 32, a);        // O_0, dims 32 g ..
    tmem_ld32(tmem_base + lane_off + 192 + g * 32, c);        // O_1, same dims
    if (qr < L) {
      uint4 *dst = reinterpret_cast<uint4 *>(out + (int64_t)(r0 + qr) * d + h * HD + g * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 8 * j + 2 * k;
          // a group that saw no key at all holds an untouched accumulator only if n_tiles == 0 (impossible);
          // a fully masked group has O_g == 0 exactly and w_g == 0
          const float x0 = __uint_as_float(a[i]) * f0 + __uint_as_float(c[i]) * f1;
          const float x1 = __uint_as_float(a[i + 1]) * f0 + __uint_as_float(c[i + 1]) * f1;
          __nv_bfloat162 v = __floats2bfloat162_rn(x0, x1);
          w[k] = *reinterpret_cast<uint32_t *>(&v);
        }
        dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
      }
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
// Persistent form of the same pipeline: 2 CTAs per SM, each walking the work items
// (query tile, head, sequence) w = blockIdx.x, blockIdx.x + gridDim.x, ...  A CTA of the one-item-per-CTA kernel
// above spends ~2.7 us in front of its first softmax round (barrier / TMEM setup, Q and K/V loads from a cold start,
// first MMA) and ~2 us behind the last one (accumulator drain, merge, stores) out of a ~19 us life at L = 1024
// (profiles/round2_fa_rr_experiment_regions.csv).  Here the TMA warp runs ahead across item boundaries (the K/V
// ring never drains; Q of the next item is fetched as soon as the last S MMA of the current one has retired), the MMA
// warp issues the next item's first scores while the softmax warps are still merging the current one, and the
// barriers keep running phase counters instead of being re-initialised.
struct Item {
  int b, h, q0, r0, L, S, c1, n_tiles;
  bool valid;
};
__device__ __forceinline__ Item make_item(int w, int nq, int n_head, int n_items, const int32_t *cu_seqlens,
                                          const int32_t *text_lens, const int32_t *seg1_lens, int mask_mode,
                                          int skip_partial) {
  Item it;
  it.valid = false;
  if (w >= n_items) return it;
  const int qt = w % nq;
  const int bh = w / nq;
  it.h = bh % n_head;
  it.b = bh / n_head;
  it.r0 = cu_seqlens[it.b];
  it.L = cu_seqlens[it.b + 1] - it.r0;
  it.q0 = qt * BQ;
  if (it.q0 >= it.L) return it;
  if (it.q0 + BQ > it.L) {
    if (skip_partial == 1) return it;
    if (skip_partial == 2 && it.L >= BQ) it.q0 = it.L - BQ;
  }
  it.S = (mask_mode != VB_MASK_FULL) ? text_lens[it.b] : 0;
  it.c1 = (mask_mode >= VB_MASK_PADDED_AR) ? seg1_lens[it.b] : 0;
  const int q_hi = min(it.q0 + BQ, it.L);
  const int kv_max = (mask_mode == VB_MASK_VALLE_AR) ? max(it.S, q_hi) : it.L;
  it.n_tiles = (kv_max + BKV - 1) / BKV;
  it.valid = it.n_tiles > 0;
  return it;
}

__global__ void __launch_bounds__(kThreads, 2)
attn_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmap, int n_head, int nq, int n_items,
                               const int32_t *__restrict__ cu_seqlens, const int32_t *__restrict__ text_lens,
                               const int32_t *__restrict__ seg1_lens, int seg1_start, int mask_mode,
                               bf16 *__restrict__ out, int skip_partial) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t *sQ = smem;
  uint8_t *sKV = sQ + kQBytes;                       // [stage][K | V]
  uint8_t *sP = sKV + kStages * kStageBytes;         // [group][128 rows x 128 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sP + kPBytes);
  uint64_t *q_full = bars;              // Q tile of the current item landed
  uint64_t *kv_full = bars + 1;         // [kStages]
  uint64_t *kv_empty = kv_full + kStages;
  uint64_t *s_full = kv_empty + kStages;   // [2] S_g ready in TMEM
  uint64_t *p_full = s_full + 2;           // [2] P_g in smem, S_g consumed (4 warp arrivals)
  uint64_t *o_done = p_full + 2;           // [2] P_g V accumulated: P_g buffer free, O_g stable
  uint64_t *q_empty = o_done + 2;          // every S MMA of the item retired: the Q tile may be overwritten
  uint64_t *o_free = q_empty + 1;          // the merge has read O_0 and O_1 (8 warp arrivals)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_free + 1);
  static_assert((1 + 2 * kStages + 8) * 8 + 8 <= 128, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = n_head * HD;

  if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, 8);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&o_done[g], 1);
    }
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
    // ===== TMA producer: runs ahead of the consumers across item boundaries =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, n_done = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const Item it = make_item(w, nq, n_head, n_items, cu_seqlens, text_lens, seg1_lens, mask_mode, skip_partial);
        if (!it.valid) continue;
        mbar_wait(q_empty, (n_done & 1) ^ 1);            // passes at once for the first item
        mbar_expect_tx(q_full, kQBytes);
        tma_load_2d(&tmap, q_full, sQ, it.h * HD, it.r0 + it.q0);
        for (int t = 0; t < it.n_tiles; ++t) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          uint8_t *dst = sKV + stage * kStageBytes;
          mbar_expect_tx(&kv_full[stage], kStageBytes);
          tma_load_2d(&tmap, &kv_full[stage], dst, d + it.h * HD, it.r0 + t * BKV);
          tma_load_2d(&tmap, &kv_full[stage], dst + kKBytes, 2 * d + it.h * HD, it.r0 + t * BKV);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ++n_done;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(BQ, 64);       // S_g = Q K_g^T : both K-major, N = 64 keys
      constexpr uint32_t idesc_o = make_idesc_bmn(BQ, HD);   // O_g += P_g V_g : A K-major, B MN-major
      constexpr int kHalfBytes = 64 * HD * 2;                // 64 keys of a K or V tile = 8 swizzle atoms
      const uint64_t qdesc = make_smem_desc(smem_u32(sQ));
      int stage = 0;
      uint32_t phase = 0, n_done = 0, round = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const Item it = make_item(w, nq, n_head, n_items, cu_seqlens, text_lens, seg1_lens, mask_mode, skip_partial);
        if (!it.valid) continue;
        const int n_tiles = it.n_tiles;
        mbar_wait(q_full, n_done & 1);
        mbar_wait(&kv_full[stage], phase);
        tcgen05_fence_after();
        for (int g = 0; g < 2; ++g) {        // S_g(0): the score buffers were consumed in the previous item's last round
          const uint64_t kdesc = make_smem_desc(smem_u32(sKV + stage * kStageBytes + g * kHalfBytes));
#pragma unroll
          for (int k = 0; k < HD / UMMA_K; ++k)
            umma_bf16(tmem_base + g * 64, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0);
          tcgen05_commit(&s_full[g]);
        }
        if (n_tiles == 1) tcgen05_commit(q_empty);
        for (int t = 0; t < n_tiles; ++t) {
          const int vstage = stage;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&p_full[g], round & 1);          // P_g(t) written, S_g(t) consumed
            tcgen05_fence_after();
            if (t + 1 < n_tiles) {                    // S_g(t+1) first: its softmax group waits on it
              if (g == 0) {
                mbar_wait(&kv_full[stage], phase);
                tcgen05_fence_after();
              }
              const uint64_t kdesc = make_smem_desc(smem_u32(sKV + stage * kStageBytes + g * kHalfBytes));
#pragma unroll
              for (int k = 0; k < HD / UMMA_K; ++k)
                umma_bf16(tmem_base + g * 64, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0);
              tcgen05_commit(&s_full[g]);
              if (g == 1 && t + 2 == n_tiles) tcgen05_commit(q_empty);   // last S MMA of the item issued
            }
            if (t == 0 && g == 0 && n_done > 0) {      // the previous item's merge has read both accumulators
              mbar_wait(o_free, (n_done - 1) & 1);
              tcgen05_fence_after();
            }
            const uint64_t pdesc = make_smem_desc(smem_u32(sP + g * kQBytes));
            const uint64_t vdesc = make_smem_desc_mn(smem_u32(sKV + vstage * kStageBytes + kKBytes + g * kHalfBytes));
#pragma unroll
            for (int k = 0; k < 64 / UMMA_K; ++k)
              umma_bf16(tmem_base + 128 + g * 64, pdesc + (uint64_t)(k * 2), vdesc + (uint64_t)(k * (2048 >> 4)),
                        idesc_o, (t > 0 || k > 0) ? 1u : 0u);
            tcgen05_commit(&o_done[g]);
          }
          tcgen05_commit(&kv_empty[vstage]);  // K(t), V(t) free once these MMAs retire
          ++round;
        }
        ++n_done;
      }
    }
    __syncwarp();
  } else {
    // ===== softmax group g: one thread per query row, keys 128 t + 64 g .. + 63 =====
    const int quarter = warp & 3;                  // TMEM lanes this warp may touch
    const int g = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + g * 64;
    const uint32_t o_addr = tmem_base + lane_off + 128 + g * 64;
    uint8_t *p_row = sP + g * kQBytes + row * 128;
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    uint32_t round = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item it = make_item(w, nq, n_head, n_items, cu_seqlens, text_lens, seg1_lens, mask_mode, skip_partial);
      if (!it.valid) continue;
      const int n_tiles = it.n_tiles, L = it.L;
      const int qr = it.q0 + row;
      const RowMask rm = make_row_mask(mask_mode, qr, L, it.S, seg1_start, it.c1);
      float m_ref = -CUDART_INF_F, l = 0.f;
      for (int t = 0; t < n_tiles; ++t, ++round) {
        const int jc0 = t * BKV + g * 64;
        if (lane == 0) mbar_wait(&s_full[g], round & 1);
        __syncwarp();
        tcgen05_fence_after();
        const bool interior = __all_sync(0xffffffffu, jc0 + 64 <= rm.lim0);
        const float mx = interior ? half_row_max<false>(s_addr, rm, jc0) : half_row_max<true>(s_addr, rm, jc0);
        const float m_cand = fmaxf(m_ref, mx * sc);
        const bool need = (m_cand - m_ref) > 8.f;     // NaN (-inf - -inf) compares false
        if (t > 0) {                                  // P_g buffer free, O_g stable
          if (lane == 0) mbar_wait(&o_done[g], (round - 1) & 1);
          __syncwarp();
          tcgen05_fence_after();
        }
        if (__any_sync(0xffffffffu, need)) {
          const float corr = (m_cand == m_ref) ? 1.f : ex2(m_ref - m_cand);
          if (t > 0) {
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
              uint32_t r[32];
              tmem_ld32(o_addr + c0, r);
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
              tmem_st32(o_addr + c0, r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          l *= corr;
          m_ref = m_cand;
        }
        const float m_use = m_ref == -CUDART_INF_F ? 0.f : m_ref;
        l += interior ? half_row_p<false>(s_addr, rm, jc0, sc, m_use, p_row, row)
                      : half_row_p<true>(s_addr, rm, jc0, sc, m_use, p_row, row);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P visible to the tensor-core proxy
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // ---- merge the two key halves: out = (O_0 w_0 + O_1 w_1) / (l_0 w_0 + l_1 w_1), w_g = 2^(m_g - max) ----
      if (lane == 0) {
        mbar_wait(&o_done[0], (round - 1) & 1);
        mbar_wait(&o_done[1], (round - 1) & 1);   // in-order retirement: every MMA of this item is done, so both P
      }                                           // buffers are free and carry the (m, l) exchange
      __syncwarp();
      tcgen05_fence_after();
      // exchange slots of query row r: the first 16 bytes of P_0's row r, which only thread (group 0, row r) writes
      // again (next item, after the second pair barrier below)
      float2 *xch = reinterpret_cast<float2 *>(sP + row * 128);
      xch[g] = make_float2(m_ref, l);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      const float2 e0 = xch[0], e1 = xch[1];
      uint32_t a[32], c[32];
      tmem_ld32(tmem_base + lane_off + 128 + g * 32, a);        // O_0, dims 32 g ..
      tmem_ld32(tmem_base + lane_off + 192 + g * 32, c);        // O_1, same dims
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");   // both partners have read the exchange slots
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);            // the next item's first P V may overwrite the accumulators
      const float mm = fmaxf(e0.x, e1.x);
      const float w0 = e0.x == -CUDART_INF_F ? 0.f : ex2(e0.x - mm);
      const float w1 = e1.x == -CUDART_INF_F ? 0.f : ex2(e1.x - mm);
      const float inv = 1.f / (e0.y * w0 + e1.y * w1);
      const float f0 = w0 * inv, f1 = w1 * inv;
      if (qr < L) {
        uint4 *dst = reinterpret_cast<uint4 *>(out + (int64_t)(it.r0 + qr) * d + it.h * HD + g * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t wv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 8 * j + 2 * k;
            const float x0 = __uint_as_float(a[i]) * f0 + __uint_as_float(c[i]) * f1;
            const float x1 = __uint_as_float(a[i + 1]) * f0 + __uint_as_float(c[i + 1]) * f1;
            __nv_bfloat162 v = __floats2bfloat162_rn(x0, x1);
            wv[k] = *reinterpret_cast<uint32_t *>(&v);
          }
          dst[j] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
      }
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

}  // namespace fa5

bool attention_tcgen05_enabled() { return getenv("VB_ATTN_MMA_SYNC") == nullptr; }

int launch_attention_tcgen05(const bf16 *qkv, int64_t M, int B, int n_head, const int32_t *cu_seqlens,
                             const int32_t *text_lens, const int32_t *seg1_lens, int seg1_start, int max_seqlen,
                             int mask_mode, bf16 *out, int skip_partial, cudaStream_t s) {
  if (M == 0 || B == 0) return VB_OK;
  const int d = n_head * fa5::HD;
  CUtensorMap tm;
  VB_TRY(tc::make_tmap(&tm, qkv, M, 3 * d, 3 * (int64_t)d, fa5::BQ));
  static PerDeviceOnce once;
  if (once.first()) {
    VB_CUDA(cudaFuncSetAttribute(fa5::attn_tcgen05_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fa5::kSmemBytes));
    VB_CUDA(cudaFuncSetAttribute(fa5::attn_tcgen05_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fa5::kSmemBytes));
  }
  const int nq = (max_seqlen + fa5::BQ - 1) / fa5::BQ;
  const int64_t n_items = (int64_t)nq * n_head * B;
  if (tune("VB_FA_PERSISTENT", 1) != 0 && n_items > 2 * sm_count() && n_items < (1ll << 31)) {
    // more than one wave: 2 CTAs per SM walk the items
    fa5::attn_tcgen05_persistent_kernel<<<2 * sm_count(), fa5::kThreads, fa5::kSmemBytes, s>>>(
        tm, n_head, nq, (int)n_items, cu_seqlens, text_lens, seg1_lens, seg1_start, mask_mode, out, skip_partial);
  } else {
    dim3 grid(nq, n_head, B);
    fa5::attn_tcgen05_pp_kernel<<<grid, fa5::kThreads, fa5::kSmemBytes, s>>>(tm, n_head, cu_seqlens, text_lens, seg1_lens,
                                                                           seg1_start, mask_mode, out, skip_partial);
  }
  VB_LAUNCH_CHECK();
  return VB_OK;
}

}  // namespace vb

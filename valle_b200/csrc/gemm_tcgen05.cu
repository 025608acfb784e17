// bf16 GEMM on the 5th-generation tensor cores (sm_100a): C[M,N] = epi(A[M,K] W[N,K]^T + bias).
//
// Hand-written tcgen05 / TMEM / TMA kernel for the dense QKV / out-proj / FFN / head projections
// of the NAR passes, the AR prefill and the training forward (F.linear at
// valle/modules/activation.py:408, valle/modules/transformer.py:332-334, valle/models/valle.py:1128).
//
// Structure (persistent, one CTA per SM, 256 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of a 128 x 64 A box and a BN x 64 W box
//               (both K-major, 128-byte swizzle) into a kStages-deep shared-memory ring, mbarrier
//               complete_tx signalling.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//               (M=128, N=BN, K=16) x 4 per stage, accumulating fp32 in TMEM; tcgen05.commit frees
//               the smem stage and, after the last k-block, publishes the accumulator.
//   warp 2      TMEM allocator (2 accumulator buffers of BN columns -> epilogue of tile i overlaps
//               the MMAs of tile i+1).
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns at a time, + bias, ReLU / residual,
//               convert, 16-byte global stores.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05_ptx.cuh"

namespace vb {

namespace tc {

constexpr int BM = 128;
constexpr int kThreads = 256;

template <int BN> struct Cfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;  // 16 KB
  static constexpr int kBBytes = BN * BK * 2;  // 32 KB / 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 2 * BM * 128;  // two 128-row x 128-byte epilogue staging boxes
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;     // 512 / 256: power of two >= 32
};

template <int BN, int kEpi, typename TC>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const float *__restrict__ bias, int M, int N, int K) {
  using cfg = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *tiles = smem;
  uint8_t *staging = smem + cfg::kStages * cfg::kStageBytes;  // [2][128 rows x 128 B], 128B-swizzled
  uint64_t *bars = reinterpret_cast<uint64_t *>(staging + cfg::kStagingBytes);
  uint64_t *full_bar = bars;                       // [kStages]
  uint64_t *empty_bar = bars + cfg::kStages;       // [kStages]
  uint64_t *tmem_full = bars + 2 * cfg::kStages;   // [2]
  uint64_t *tmem_empty = tmem_full + 2;            // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM, n_tiles = N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = K / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(cfg::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / n_tiles, n_blk = t % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t *a_dst = tiles + stage * cfg::kStageBytes;
          uint8_t *b_dst = a_dst + cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], cfg::kStageBytes);
          tma_load_2d(&tmap_a, &full_bar[stage], a_dst, kb * BK, m_blk * BM);
          tma_load_2d(&tmap_b, &full_bar[stage], b_dst, kb * BK, n_blk * BN);
          if (++stage == cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * cfg::kStageBytes);
          const uint32_t b_addr = a_addr + cfg::kABytes;
          const uint64_t adesc = make_smem_desc(a_addr);
          const uint64_t bdesc = make_smem_desc(b_addr);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance along K inside the 128-byte swizzle atom: +32 bytes (>>4 = 2) per UMMA_K
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          tcgen05_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
          if (++stage == cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tcgen05_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers (+bias, ReLU) -> 128B-swizzled staging box -> TMA store =====
    // fp32 residual outputs use cp.reduce.async.bulk (.add): the residual add of transformer.py:297-302
    // happens in the memory system, x is never read by the SM.
    constexpr int kColsPerUnit = 128 / (int)sizeof(TC);  // 32 fp32 or 64 bf16 columns = one 128-byte row
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row_l = q * 32 + lane;
    const bool issuer = (warp == 4 && lane == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    int unit = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / n_tiles, n_blk = t % n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const float *brow = bias ? bias + n_blk * BN : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += kColsPerUnit, ++unit) {
        uint8_t *box = staging + (unit & 1) * (BM * 128);
        // the TMA store issued two units ago (same box) must have finished READING shared memory
        if (issuer) tma_store_wait_read<1>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint8_t *srow = box + row_l * 128;
#pragma unroll
        for (int h = 0; h < kColsPerUnit / 32; ++h) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0 + h * 32), r);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = __uint_as_float(r[i]);
            if (brow) v[i] += __ldg(brow + c0 + h * 32 + i);
            if constexpr (kEpi == VB_EPI_RELU) v[i] = fmaxf(v[i], 0.f);
          }
          if constexpr (sizeof(TC) == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j)  // 8 chunks of 4 floats
              *reinterpret_cast<float4 *>(srow + ((j ^ (row_l & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 4 chunks of 8 bf16 (this half of the 64-column unit)
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
              __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
              __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
              const int chunk = h * 4 + j;
              *reinterpret_cast<uint4 *>(srow + ((chunk ^ (row_l & 7)) << 4)) =
                  make_uint4(*reinterpret_cast<uint32_t *>(&p0), *reinterpret_cast<uint32_t *>(&p1),
                             *reinterpret_cast<uint32_t *>(&p2), *reinterpret_cast<uint32_t *>(&p3));
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging visible to the TMA engine
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          if constexpr (kEpi == VB_EPI_RESIDUAL)
            tma_reduce_add_2d(&tmap_c, box, n_blk * BN + c0, m_blk * BM);
          else
            tma_store_2d(&tmap_c, box, n_blk * BN + c0, m_blk * BM);
          tma_store_commit();
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (issuer) tma_store_wait_read<0>();  // shared memory must outlive the last store's reads
  }
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(cfg::kTmemCols));
  }
}

template <int BN, int kEpi, typename TC>
static int launch_t(const CUtensorMap &ta, const CUtensorMap &tb, const float *bias, TC *C, int64_t ldc, int M,
                    int N, int K, cudaStream_t s) {
  CUtensorMap tcm;
  VB_TRY(tc::make_tmap_2d(&tcm, C, M, N, ldc, BM, 128 / (int)sizeof(TC), sizeof(TC) == 4));
  using cfg = Cfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, kEpi, TC>;
  static PerDeviceOnce once;  // per template instantiation and device
  if (once.first()) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::kSmemBytes));
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, kThreads, cfg::kSmemBytes, s>>>(ta, tb, tcm, bias, M, N, K);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

template <int BN>
static int launch_bn(const CUtensorMap &ta, const CUtensorMap &tb, const float *bias, void *C, int c_dtype,
                     int64_t ldc, int M, int N, int K, int epi, cudaStream_t s) {
  if (epi == VB_EPI_RESIDUAL) return launch_t<BN, VB_EPI_RESIDUAL, float>(ta, tb, bias, (float *)C, ldc, M, N, K, s);
  if (epi == VB_EPI_RELU) {
    if (c_dtype == VB_BF16) return launch_t<BN, VB_EPI_RELU, bf16>(ta, tb, bias, (bf16 *)C, ldc, M, N, K, s);
    return launch_t<BN, VB_EPI_RELU, float>(ta, tb, bias, (float *)C, ldc, M, N, K, s);
  }
  if (c_dtype == VB_BF16) return launch_t<BN, VB_EPI_NONE, bf16>(ta, tb, bias, (bf16 *)C, ldc, M, N, K, s);
  return launch_t<BN, VB_EPI_NONE, float>(ta, tb, bias, (float *)C, ldc, M, N, K, s);
}

}  // namespace tc

bool tcgen05_gemm_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc) {
  return M >= 1 && M < (1LL << 31) && N % 128 == 0 && K % tc::BK == 0 && lda % 8 == 0 && ldc % 8 == 0 &&
         getenv("VB_DISABLE_TCGEN05") == nullptr;
}

int launch_gemm_tcgen05(const bf16 *A, int64_t lda, const bf16 *W, const float *bias, void *C, int c_dtype,
                        int64_t ldc, int64_t M, int N, int K, int epi, cudaStream_t s) {
  VB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "tcgen05 gemm: operands must be 16-byte aligned");
  const int m_tiles = (int)((M + tc::BM - 1) / tc::BM);
  const bool wide = (N % 256 == 0) && (m_tiles * (N / 256) >= sm_count());
  const int BN = wide ? 256 : 128;
  CUtensorMap ta, tb;
  VB_TRY(tc::make_tmap(&ta, A, M, K, lda, tc::BM));
  VB_TRY(tc::make_tmap(&tb, W, N, K, K, BN));
  if (BN == 256) return tc::launch_bn<256>(ta, tb, bias, C, c_dtype, ldc, (int)M, N, K, epi, s);
  return tc::launch_bn<128>(ta, tb, bias, C, c_dtype, ldc, (int)M, N, K, epi, s);
}

}  // namespace vb

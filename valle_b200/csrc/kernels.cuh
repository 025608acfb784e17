// Internal launcher declarations shared by the translation units of libvalle_b200.so.
#pragma once
#include "common.cuh"

namespace vb {

struct LnParams {
  const float *gamma, *beta, *ada_wb;  // ada_wb: NULL or [2d] (weight | bias)
  float eps;
};

struct QkvScatter {
  int d, head_dim;
  float *q;  // [B, d] fp32
  void *kcache, *vcache;  // this layer's cache base
  int64_t cache_seq_stride;
  int cache_cap;
  const int32_t *text_len, *prompt_len, *n_gen;
};

// gemm_simt.cu
int launch_gemm_simt(const void *A, int a_dtype, int64_t lda, const void *W, const float *bias, void *C,
                     int c_dtype, int64_t ldc, int64_t M, int N, int K, int epi, cudaStream_t s);
int launch_gemv(const float *x, int64_t ldx, int B, const void *W, int w_dtype, const float *bias,
                int N, int K, float *out, int64_t ldo, const LnParams *ln, int epi_mode,
                const QkvScatter *qkv, cudaStream_t s);

// gemm_tcgen05.cu
bool tcgen05_gemm_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc);
int launch_gemm_tcgen05(const bf16 *A, int64_t lda, const bf16 *W, const float *bias, void *C,
                        int c_dtype, int64_t ldc, int64_t M, int N, int K, int epi, cudaStream_t s);

// gemm_decode.cu (swap-AB split-K tcgen05 projections for B <= 64 decode rows, bf16)
enum { DG_F32 = 0, DG_RESIDUAL = 1, DG_RELU_BF16 = 2, DG_QKV = 3 };
size_t gemm_decode_workspace();
int launch_gemm_decode(const bf16 *act, int B, int64_t ld_act, const bf16 *W, int N, int K, int force_splits,
                       const float *bias, int mode, float *out_f32, bf16 *out_bf16, int64_t ld_out,
                       const QkvScatter *qkv, float *partials, size_t partial_bytes, int *out_splits, int *out_ldp,
                       bool pdl, cudaStream_t s);

// attention.cu
int launch_attention_varlen(const void *qkv, int dtype, int64_t M, int B, int n_head, int head_dim,
                            const int32_t *cu_seqlens, const int32_t *text_lens, int max_seqlen,
                            int mask_mode, void *out, void *kcache, void *vcache,
                            int64_t cache_seq_stride, int cache_cap, cudaStream_t s);
// attention_mma.cu (bf16 tensor-core flash attention)
int launch_attention_mma(const bf16 *qkv, int64_t M, int B, int n_head, const int32_t *cu_seqlens,
                         const int32_t *text_lens, int max_seqlen, int mask_mode, bf16 *out, bf16 *kcache,
                         bf16 *vcache, int64_t cache_seq_stride, int cache_cap, cudaStream_t s);
size_t attn_decode_workspace(int B, int n_head, int head_dim, int cache_cap);
int launch_attn_decode(const float *q, const float *qkv_part, int qkv_splits, int qkv_ldp, const float *qkv_bias,
                       int B, int n_head, int head_dim, void *kcache, void *vcache, int dtype,
                       int64_t cache_seq_stride, int cache_cap, const int32_t *text_len, const int32_t *prompt_len,
                       const int32_t *n_gen, float *out, void *out16, void *workspace, bool pdl, cudaStream_t s);

// decode_fused.cu
int launch_relu_reduce(const float *partials, int splits, int ldp, const float *bias, int B, int N, bf16 *out16,
                       int64_t ldo, bool pdl, cudaStream_t s);
int launch_ln_reduce(float *x, int64_t ldx, int B, int d, const float *partials, int splits, int ldp,
                     const float *bias, const float *gamma, const float *beta, float eps, bf16 *out16, bool pdl,
                     cudaStream_t s);

// sample.cu
int launch_ar_sample(float *logits, int64_t ld_logits, const float *partials, int splits, int ldp,
                     const vb_ar_head *head, vb_ar_state *st, int d, const int64_t *forced, int reduce_only, bool pdl,
                     cudaStream_t s);

}  // namespace vb

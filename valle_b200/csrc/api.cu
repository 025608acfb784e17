// C-ABI entry points and host-side orchestration (layer loops) of libvalle_b200.so.
// See include/valle_b200.h for the contract of every function.
#include <stdarg.h>
#include <string.h>

#include <new>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace vb {

static thread_local char g_err[1024] = "";
static int64_t g_launches = 0;  // process-wide counter (relaxed; bench reads it when idle)

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED); }

// Tuning knobs: vb_tune_set() overrides > environment variable of the same name > built-in default.  Read at
// launch time (a captured CUDA graph keeps the values it was captured with).
namespace {
struct TuneEntry {
  char name[48];
  int value;
};
TuneEntry g_tune[32];
int g_n_tune = 0;
}  // namespace
int tune(const char *name, int dflt) {
  for (int i = 0; i < g_n_tune; ++i)
    if (strcmp(g_tune[i].name, name) == 0) return g_tune[i].value;
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

#ifdef VB_TRACE
static trace_bind_fn g_trace_binders[32];
static int g_n_trace_binders = 0;
void trace_register(trace_bind_fn f) {
  if (g_n_trace_binders < 32) g_trace_binders[g_n_trace_binders++] = f;
}
#endif

}  // namespace vb

using namespace vb;

struct vb_decoder {
  vb_decoder_desc desc;
  vb_layer_params *layers;  // owned host copy
  vb_ln_fold *fold_qkv = nullptr, *fold_ffn1 = nullptr;  // owned host copies [n_layer] or NULL (vb_decoder_set_decode_fold)
};

VB_API int vb_abi_version(void) { return VB_ABI_VERSION; }
VB_API const char *vb_last_error(void) { return g_err; }
VB_API int64_t vb_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

VB_API int vb_tune_set(const char *name, int value) {
  VB_CHECK_ARG(name && strlen(name) < sizeof(g_tune[0].name), "vb_tune_set: bad name");
  for (int i = 0; i < g_n_tune; ++i)
    if (strcmp(g_tune[i].name, name) == 0) {
      g_tune[i].value = value;
      return VB_OK;
    }
  VB_CHECK_ARG(g_n_tune < 32, "vb_tune_set: table full");
  strcpy(g_tune[g_n_tune].name, name);
  g_tune[g_n_tune++].value = value;
  return VB_OK;
}

VB_API int vb_trace_bind(unsigned long long *buf, unsigned int *counter, unsigned int cap) {
#ifdef VB_TRACE
  for (int i = 0; i < g_n_trace_binders; ++i)
    if (g_trace_binders[i](buf, counter, cap) != 0) {
      set_error("vb_trace_bind: cudaMemcpyToSymbol failed");
      return VB_ERR_CUDA;
    }
  return VB_OK;
#else
  (void)buf; (void)counter; (void)cap;
  set_error("vb_trace_bind: not a profiling build (compile with -DVB_TRACE: python -m valle_b200.build --trace)");
  return VB_ERR_UNSUPPORTED;
#endif
}

VB_API int vb_linear(const void *A, int a_dtype, int64_t lda, const void *W, int w_dtype,
                         const float *bias, void *C, int c_dtype, int64_t ldc, int64_t M, int N, int K,
                         int epilogue, void *workspace, size_t workspace_bytes, vb_stream_t stream) {
  (void)workspace;
  (void)workspace_bytes;
  VB_CHECK_ARG(a_dtype == w_dtype, "vb_linear: a_dtype (%d) must equal w_dtype (%d)", a_dtype, w_dtype);
  VB_CHECK_ARG(epilogue >= VB_EPI_NONE && epilogue <= VB_EPI_RESIDUAL, "vb_linear: bad epilogue %d", epilogue);
  VB_CHECK_ARG(M >= 0 && N > 0 && K > 0, "vb_linear: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  cudaStream_t s = (cudaStream_t)stream;
  if (a_dtype == VB_BF16 && tcgen05_gemm_supported(M, N, K, lda, ldc))
    return launch_gemm_tcgen05((const bf16 *)A, lda, (const bf16 *)W, bias, C, c_dtype, ldc, M, N, K,
                               epilogue, s);
  return launch_gemm_simt(A, a_dtype, lda, W, bias, C, c_dtype, ldc, M, N, K, epilogue, s);
}

VB_API int vb_decoder_create(const vb_decoder_desc *desc, vb_decoder_t *out) {
  VB_CHECK_ARG(desc && out, "vb_decoder_create: null argument");
  VB_CHECK_ARG(desc->n_layer > 0 && desc->n_head > 0 && desc->d_model % desc->n_head == 0,
               "vb_decoder_create: bad geometry d=%d H=%d L=%d", desc->d_model, desc->n_head, desc->n_layer);
  VB_CHECK_ARG(desc->d_model / desc->n_head == 64, "vb_decoder_create: head_dim must be 64 (got %d)",
               desc->d_model / desc->n_head);
  VB_CHECK_ARG(desc->d_model % 256 == 0 && desc->d_ff % 256 == 0, "vb_decoder_create: d_model and d_ff must be multiples of 256");
  VB_CHECK_ARG(desc->wdtype == VB_F32 || desc->wdtype == VB_BF16, "vb_decoder_create: bad wdtype");
  vb_decoder *d = new (std::nothrow) vb_decoder;
  if (!d) {
    set_error("vb_decoder_create: out of host memory");
    return VB_ERR_ARG;
  }
  d->desc = *desc;
  d->layers = new (std::nothrow) vb_layer_params[desc->n_layer];
  if (!d->layers) {
    delete d;
    set_error("vb_decoder_create: out of host memory");
    return VB_ERR_ARG;
  }
  memcpy(d->layers, desc->layers, sizeof(vb_layer_params) * desc->n_layer);
  d->desc.layers = d->layers;
  *out = d;
  return VB_OK;
}

VB_API void vb_decoder_destroy(vb_decoder_t dec) {
  if (!dec) return;
  delete[] dec->layers;
  delete[] dec->fold_qkv;
  delete[] dec->fold_ffn1;
  delete dec;
}

VB_API int vb_ln_fold_build(const void *W, int N, int K, const float *gamma, const float *beta, const float *bias,
                            void *wf, float *c, float *dvec, vb_stream_t stream) {
  VB_CHECK_ARG(W && gamma && beta && wf && c && dvec && N > 0 && K > 0, "vb_ln_fold_build: null argument / bad shape");
  return launch_ln_fold((const bf16 *)W, N, K, gamma, beta, bias, (bf16 *)wf, c, dvec, (cudaStream_t)stream);
}

VB_API int vb_decoder_set_decode_fold(vb_decoder_t dec, const vb_ln_fold *qkv, const vb_ln_fold *ffn1) {
  VB_CHECK_ARG(dec, "vb_decoder_set_decode_fold: null decoder");
  delete[] dec->fold_qkv;
  delete[] dec->fold_ffn1;
  dec->fold_qkv = dec->fold_ffn1 = nullptr;
  if (!qkv && !ffn1) return VB_OK;
  VB_CHECK_ARG(qkv && ffn1, "vb_decoder_set_decode_fold: both arrays or neither");
  VB_CHECK_ARG(dec->desc.wdtype == VB_BF16, "vb_decoder_set_decode_fold: bf16 decoders only");
  const int n = dec->desc.n_layer;
  for (int l = 0; l < n; ++l)
    VB_CHECK_ARG(qkv[l].wf && qkv[l].c && qkv[l].dvec && ffn1[l].wf && ffn1[l].c && ffn1[l].dvec,
                 "vb_decoder_set_decode_fold: layer %d has a null pointer", l);
  dec->fold_qkv = new (std::nothrow) vb_ln_fold[n];
  dec->fold_ffn1 = new (std::nothrow) vb_ln_fold[n];
  if (!dec->fold_qkv || !dec->fold_ffn1) {
    delete[] dec->fold_qkv;
    delete[] dec->fold_ffn1;
    dec->fold_qkv = dec->fold_ffn1 = nullptr;
    set_error("vb_decoder_set_decode_fold: out of host memory");
    return VB_ERR_ARG;
  }
  memcpy(dec->fold_qkv, qkv, sizeof(vb_ln_fold) * n);
  memcpy(dec->fold_ffn1, ffn1, sizeof(vb_ln_fold) * n);
  return VB_OK;
}

static size_t elem_size(int dtype) { return dtype == VB_BF16 ? 2 : 4; }

VB_API size_t vb_decoder_forward_workspace(const vb_decoder_desc *desc, int64_t M) {
  const size_t ts = elem_size(desc->wdtype);
  const size_t d = desc->d_model, dff = desc->d_ff;
  const size_t Mp = align_up((size_t)M, 128);
  return Mp * (d + 3 * d + d + dff) * ts + 4 * 256;
}

VB_API int vb_decoder_forward(vb_decoder_t dec, float *x, int64_t M, int B, const int32_t *cu_seqlens,
                                  const int32_t *text_lens, const int32_t *seg1_lens, int seg1_start,
                                  int max_seqlen, int mask_mode,
                                  const float *ada_wb, void *kcache, void *vcache,
                                  int64_t cache_layer_stride, int64_t cache_seq_stride, int cache_cap,
                                  void *workspace, size_t workspace_bytes, vb_stream_t stream) {
  VB_CHECK_ARG(dec && x && cu_seqlens, "vb_decoder_forward: null argument");
  const vb_decoder_desc &D = dec->desc;
  VB_CHECK_ARG(workspace_bytes >= vb_decoder_forward_workspace(&D, M),
               "vb_decoder_forward: workspace too small (%zu < %zu)", workspace_bytes,
               vb_decoder_forward_workspace(&D, M));
  if (M == 0) return VB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int d = D.d_model, dff = D.d_ff, dt = D.wdtype;
  const size_t ts = elem_size(dt);
  const size_t Mp = align_up((size_t)M, 128);
  char *ws = (char *)workspace;
  void *xn = ws;   ws += align_up(Mp * d * ts, 256);
  void *qkv = ws;  ws += align_up(Mp * 3 * d * ts, 256);
  void *att = ws;  ws += align_up(Mp * d * ts, 256);
  void *hb = ws;
  for (int l = 0; l < D.n_layer; ++l) {
    const vb_layer_params &P = dec->layers[l];
    const float *ada1 = ada_wb ? ada_wb + (size_t)(2 * l) * 2 * d : nullptr;
    const float *ada2 = ada_wb ? ada_wb + (size_t)(2 * l + 1) * 2 * d : nullptr;
    VB_TRY(vb_layernorm(x, d, nullptr, M, d, P.norm1_w, P.norm1_b, ada1, 1e-5f, xn, dt, stream));
    VB_TRY(vb_linear(xn, dt, d, P.in_proj_w, dt, P.in_proj_b, qkv, dt, 3 * d, M, 3 * d, d, VB_EPI_NONE,
                     nullptr, 0, stream));
    void *kc = kcache ? (char *)kcache + (size_t)l * cache_layer_stride * ts : nullptr;
    void *vc = vcache ? (char *)vcache + (size_t)l * cache_layer_stride * ts : nullptr;
    VB_TRY(launch_attention_varlen(qkv, dt, M, B, D.n_head, d / D.n_head, cu_seqlens, text_lens, seg1_lens, seg1_start, max_seqlen,
                                   mask_mode, att, kc, vc, cache_seq_stride, cache_cap, nullptr, 0, s));
    VB_TRY(vb_linear(att, dt, d, P.out_proj_w, dt, P.out_proj_b, x, VB_F32, d, M, d, d, VB_EPI_RESIDUAL,
                     nullptr, 0, stream));
    VB_TRY(vb_layernorm(x, d, nullptr, M, d, P.norm2_w, P.norm2_b, ada2, 1e-5f, xn, dt, stream));
    VB_TRY(vb_linear(xn, dt, d, P.lin1_w, dt, P.lin1_b, hb, dt, dff, M, dff, d, VB_EPI_RELU, nullptr, 0,
                     stream));
    VB_TRY(vb_linear(hb, dt, dff, P.lin2_w, dt, P.lin2_b, x, VB_F32, d, M, d, dff, VB_EPI_RESIDUAL, nullptr,
                     0, stream));
  }
  return VB_OK;
}

// ------------------------------------------------------------------------------------------
// Training: forward that keeps the activations, and the backward pass of the stack
// ------------------------------------------------------------------------------------------
namespace {
struct LayerSave {
  float *x_in, *x_mid;          // fp32 [M, d] residual stream at the layer input / after the attention block
  void *xn1, *qkv, *att, *xn2, *hb;  // storage dtype: LN1 out, q|k|v, attention out, LN2 out, FFN hidden (post-ReLU)
};
size_t layer_save_bytes(const vb_decoder_desc &D, int64_t M) {
  const size_t ts = elem_size(D.wdtype), d = D.d_model, dff = D.d_ff, Mp = align_up((size_t)M, 128);
  return 2 * align_up(Mp * d * 4, 256) + 3 * align_up(Mp * d * ts, 256) + align_up(Mp * 3 * d * ts, 256) +
         align_up(Mp * dff * ts, 256);
}
LayerSave carve_layer_save(const vb_decoder_desc &D, int64_t M, char *base) {
  const size_t ts = elem_size(D.wdtype), d = D.d_model, dff = D.d_ff, Mp = align_up((size_t)M, 128);
  LayerSave s{};
  char *p = base;
  auto take = [&](size_t n) {
    char *r = p;
    p += align_up(n, 256);
    return r;
  };
  s.x_in = (float *)take(Mp * d * 4);
  s.x_mid = (float *)take(Mp * d * 4);
  s.xn1 = take(Mp * d * ts);
  s.att = take(Mp * d * ts);
  s.xn2 = take(Mp * d * ts);
  s.qkv = take(Mp * 3 * d * ts);
  s.hb = take(Mp * dff * ts);
  return s;
}
}  // namespace

VB_API size_t vb_decoder_train_save_bytes(const vb_decoder_desc *desc, int64_t M) {
  // + one fp32 [M, d] scratch behind the layers: the sub-layer output that dropout scales before the residual add
  return (size_t)desc->n_layer * layer_save_bytes(*desc, M) + align_up((size_t)M * desc->d_model * 4, 256) + 256;
}

VB_API int vb_decoder_forward_train(vb_decoder_t dec, float *x, int64_t M, int B, const int32_t *cu_seqlens,
                                    const int32_t *text_lens, const int32_t *seg1_lens, int seg1_start, int max_seqlen,
                                    int mask_mode, const float *ada_wb, void *save, size_t save_bytes,
                                    float dropout_p, uint64_t dropout_seed, vb_stream_t stream) {
  VB_CHECK_ARG(dec && x && cu_seqlens && save, "vb_decoder_forward_train: null argument");
  VB_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "vb_decoder_forward_train: dropout_p=%g not in [0, 1)", (double)dropout_p);
  const vb_decoder_desc &D = dec->desc;
  VB_CHECK_ARG(save_bytes >= vb_decoder_train_save_bytes(&D, M), "vb_decoder_forward_train: save buffer too small");
  if (M == 0) return VB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int d = D.d_model, dff = D.d_ff, dt = D.wdtype;
  const size_t per_layer = layer_save_bytes(D, M);
  const bool drop = dropout_p > 0.f;
  float *sub = (float *)((char *)save + (size_t)D.n_layer * per_layer);   // sub-layer output ahead of its dropout
  for (int l = 0; l < D.n_layer; ++l) {
    const vb_layer_params &P = dec->layers[l];
    LayerSave sv = carve_layer_save(D, M, (char *)save + (size_t)l * per_layer);
    const float *ada1 = ada_wb ? ada_wb + (size_t)(2 * l) * 2 * d : nullptr;
    const float *ada2 = ada_wb ? ada_wb + (size_t)(2 * l + 1) * 2 * d : nullptr;
    VB_CUDA(cudaMemcpyAsync(sv.x_in, x, (size_t)M * d * 4, cudaMemcpyDeviceToDevice, s));
    VB_TRY(vb_layernorm(x, d, nullptr, M, d, P.norm1_w, P.norm1_b, ada1, 1e-5f, sv.xn1, dt, stream));
    VB_TRY(vb_linear(sv.xn1, dt, d, P.in_proj_w, dt, P.in_proj_b, sv.qkv, dt, 3 * d, M, 3 * d, d, VB_EPI_NONE, nullptr, 0,
                     stream));
    // training-mode dropout (p > 0): attention probabilities (activation.py:199 `dropout=`), dropout1 / dropout2 on
    // the sub-layer outputs and `dropout` on the FFN hidden (transformer.py:329,333-334); masks from the stateless
    // hash of kernels.cuh, site streams (l << 2) | {0, 1, 2, 3}, regenerated by vb_decoder_backward
    const DropCfg dc_attn = make_drop(dropout_p, dropout_seed, (uint32_t)(l << 2) | 0u);
    VB_TRY(launch_attention_varlen(sv.qkv, dt, M, B, D.n_head, d / D.n_head, cu_seqlens, text_lens, seg1_lens, seg1_start,
                                   max_seqlen, mask_mode, sv.att, nullptr, nullptr, 0, 0, nullptr, 0, s, &dc_attn));
    if (drop) {
      VB_TRY(vb_linear(sv.att, dt, d, P.out_proj_w, dt, P.out_proj_b, sub, VB_F32, d, M, d, d, VB_EPI_NONE, nullptr, 0,
                       stream));
      VB_TRY(launch_dropout_add(x, sub, (int64_t)M * d, make_drop(dropout_p, dropout_seed, (uint32_t)(l << 2) | 1u), s));
    } else {
      VB_TRY(vb_linear(sv.att, dt, d, P.out_proj_w, dt, P.out_proj_b, x, VB_F32, d, M, d, d, VB_EPI_RESIDUAL, nullptr, 0,
                       stream));
    }
    VB_CUDA(cudaMemcpyAsync(sv.x_mid, x, (size_t)M * d * 4, cudaMemcpyDeviceToDevice, s));
    VB_TRY(vb_layernorm(x, d, nullptr, M, d, P.norm2_w, P.norm2_b, ada2, 1e-5f, sv.xn2, dt, stream));
    VB_TRY(vb_linear(sv.xn2, dt, d, P.lin1_w, dt, P.lin1_b, sv.hb, dt, dff, M, dff, d, VB_EPI_RELU, nullptr, 0, stream));
    if (drop) {
      VB_TRY(launch_dropout(sv.hb, sv.hb, dt, (int64_t)M * dff, make_drop(dropout_p, dropout_seed, (uint32_t)(l << 2) | 2u), s));
      VB_TRY(vb_linear(sv.hb, dt, dff, P.lin2_w, dt, P.lin2_b, sub, VB_F32, d, M, d, dff, VB_EPI_NONE, nullptr, 0, stream));
      VB_TRY(launch_dropout_add(x, sub, (int64_t)M * d, make_drop(dropout_p, dropout_seed, (uint32_t)(l << 2) | 3u), s));
    } else {
      VB_TRY(vb_linear(sv.hb, dt, dff, P.lin2_w, dt, P.lin2_b, x, VB_F32, d, M, d, dff, VB_EPI_RESIDUAL, nullptr, 0, stream));
    }
  }
  return VB_OK;
}

VB_API size_t vb_decoder_backward_workspace(const vb_decoder_desc *desc, int64_t M) {
  const size_t ts = elem_size(desc->wdtype), d = desc->d_model, dff = desc->d_ff, Mp = align_up((size_t)M, 128);
  size_t n = 0;
  n += align_up(Mp * d * ts, 256);        // dx in the storage dtype
  n += align_up(Mp * d * ts, 256);        // dy: dx through the mask of a sub-layer's output dropout
  n += align_up(Mp * dff * ts, 256);      // dh
  n += align_up(Mp * d * 4, 256);         // dc / da (fp32)
  n += align_up(Mp * d * ts, 256);        // do
  n += align_up(Mp * 3 * d * ts, 256);    // dqkv
  n += align_up(vb_attention_backward_workspace(M, desc->n_head), 256);
  n += align_up(vb_linear_backward_workspace(desc->wdtype, M, (int)std::max(dff, 3 * d), (int)std::max(dff, d)), 256);
  return n + 256;
}

VB_API int vb_decoder_backward(vb_decoder_t dec, float *dx, int64_t M, int B, const int32_t *cu_seqlens,
                               const int32_t *text_lens, const int32_t *seg1_lens, int seg1_start, int max_seqlen,
                               int mask_mode, const float *ada_wb, float *dada_wb, const void *save,
                               const vb_layer_wt *wt, const vb_layer_grads *grads, void *workspace,
                               size_t workspace_bytes, float dropout_p, uint64_t dropout_seed, vb_stream_t stream) {
  VB_CHECK_ARG(dec && dx && cu_seqlens && save && wt && grads, "vb_decoder_backward: null argument");
  VB_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "vb_decoder_backward: dropout_p=%g not in [0, 1)", (double)dropout_p);
  const vb_decoder_desc &D = dec->desc;
  VB_CHECK_ARG(workspace_bytes >= vb_decoder_backward_workspace(&D, M), "vb_decoder_backward: workspace too small");
  VB_CHECK_ARG(!ada_wb || dada_wb, "vb_decoder_backward: AdaLN stack needs dada_wb");
  if (M == 0) return VB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int d = D.d_model, dff = D.d_ff, dt = D.wdtype;
  const size_t ts = elem_size(dt), Mp = align_up((size_t)M, 128);
  char *p = (char *)workspace;
  auto take = [&](size_t n) {
    char *r = p;
    p += align_up(n, 256);
    return r;
  };
  void *dx_dt = take(Mp * d * ts);
  void *dy = take(Mp * d * ts);
  void *dh = take(Mp * dff * ts);
  const bool drop = dropout_p > 0.f;
  const float inv_keep = drop ? 1.f / (1.f - dropout_p) : 1.f;
  float *dn = (float *)take(Mp * d * 4);
  void *dO = take(Mp * d * ts);
  void *dqkv = take(Mp * 3 * d * ts);
  const size_t attn_ws_bytes = vb_attention_backward_workspace(M, D.n_head);
  void *attn_ws = take(attn_ws_bytes);
  const size_t lin_ws_bytes = vb_linear_backward_workspace(dt, M, std::max(dff, 3 * d), std::max(dff, d));
  void *lin_ws = take(lin_ws_bytes);
  const size_t per_layer = layer_save_bytes(D, M);
  VB_TRY(launch_cast_from_f32(dx, dx_dt, dt, (int64_t)M * d, s));
  for (int l = D.n_layer - 1; l >= 0; --l) {
    const vb_layer_params &P = dec->layers[l];
    const vb_layer_grads &G = grads[l];
    const vb_layer_wt &T = wt[l];
    LayerSave sv = carve_layer_save(D, M, (char *)const_cast<void *>(save) + (size_t)l * per_layer);
    const float *ada1 = ada_wb ? ada_wb + (size_t)(2 * l) * 2 * d : nullptr;
    const float *ada2 = ada_wb ? ada_wb + (size_t)(2 * l + 1) * 2 * d : nullptr;
    float *dada1 = dada_wb ? dada_wb + (size_t)(2 * l) * 2 * d : nullptr;
    float *dada2 = dada_wb ? dada_wb + (size_t)(2 * l + 1) * 2 * d : nullptr;
    // ---- FFN: x2 = x1 + relu(LN2(x1) W1^T + b1) W2^T + b2 (transformer.py:332-334) ----
    // (with dropout: the saved hidden is post-dropout -- zero where dropped -- and dx reaches the sub-layer output
    //  through the mask of dropout2)
    const void *dy2 = dx_dt;
    if (drop) {
      VB_TRY(launch_dropout(dx_dt, dy, dt, (int64_t)M * d, make_drop(dropout_p, dropout_seed, (uint32_t)(l << 2) | 3u), s));
      dy2 = dy;
    }
    VB_TRY(vb_linear_backward(sv.hb, dt, dff, T.lin2_wt, dy2, d, dh, dt, dff, VB_EPI_NONE, G.lin2_w, G.lin2_b, M, d, dff,
                              lin_ws, lin_ws_bytes, stream));
    VB_TRY(launch_relu_bwd(dh, sv.hb, dt, (int64_t)M * dff, inv_keep, s));
    VB_TRY(vb_linear_backward(sv.xn2, dt, d, T.lin1_wt, dh, dff, dn, VB_F32, d, VB_EPI_NONE, G.lin1_w, G.lin1_b, M, dff, d,
                              lin_ws, lin_ws_bytes, stream));
    VB_TRY(vb_layernorm_backward(sv.x_mid, d, nullptr, M, d, P.norm2_w, P.norm2_b, ada2, 1e-5f, dn, d, dx, d, dx_dt, dt,
                                 G.norm2_w, G.norm2_b, dada2, stream));
    // ---- attention block: x1 = x + Attn(LN1(x) Win^T + bin) Wo^T + bo (transformer.py:315-330) ----
    const void *dy1 = dx_dt;
    if (drop) {
      VB_TRY(launch_dropout(dx_dt, dy, dt, (int64_t)M * d, make_drop(dropout_p, dropout_seed, (uint32_t)(l << 2) | 1u), s));
      dy1 = dy;
    }
    VB_TRY(vb_linear_backward(sv.att, dt, d, T.out_proj_wt, dy1, d, dO, dt, d, VB_EPI_NONE, G.out_proj_w, G.out_proj_b, M, d,
                              d, lin_ws, lin_ws_bytes, stream));
    const DropCfg dc_attn = make_drop(dropout_p, dropout_seed, (uint32_t)(l << 2) | 0u);
    VB_TRY(attention_backward(sv.qkv, sv.att, dO, dt, M, B, D.n_head, d / D.n_head, cu_seqlens, text_lens, seg1_lens,
                              seg1_start, max_seqlen, mask_mode, dqkv, attn_ws, attn_ws_bytes, &dc_attn, s));
    VB_TRY(vb_linear_backward(sv.xn1, dt, d, T.in_proj_wt, dqkv, 3 * d, dn, VB_F32, d, VB_EPI_NONE, G.in_proj_w, G.in_proj_b,
                              M, 3 * d, d, lin_ws, lin_ws_bytes, stream));
    VB_TRY(vb_layernorm_backward(sv.x_in, d, nullptr, M, d, P.norm1_w, P.norm1_b, ada1, 1e-5f, dn, d, dx, d, dx_dt, dt,
                                 G.norm1_w, G.norm1_b, dada1, stream));
  }
  return VB_OK;
}

// ------------------------------------------------------------------------------------------
// AR decode
// ------------------------------------------------------------------------------------------
namespace {
struct StepWs {
  float *q, *att, *hb;
  void *attn_ws;
  bf16 *xn16, *att16, *hb16;
  void *gemm_ws;
  float *stats;  // moments of the folded-LayerNorm projections: [kMaxForcedSplits][64][2]
  size_t gemm_ws_bytes;
  size_t total;
};
StepWs carve_step_ws(const vb_decoder_desc &D, int B, int cache_cap, void *base) {
  const size_t d = D.d_model, dff = D.d_ff;
  StepWs w{};
  char *p = (char *)base;
  auto take = [&](size_t n) {
    char *r = p;
    p += align_up(n, 256);
    return r;
  };
  w.q = (float *)take((size_t)B * d * 4);
  w.att = (float *)take((size_t)B * d * 4);
  w.hb = (float *)take((size_t)B * dff * 4);
  w.attn_ws = take(attn_decode_workspace(B, D.n_head, (int)(d / D.n_head), cache_cap));
  w.xn16 = (bf16 *)take((size_t)64 * d * 2);
  w.att16 = (bf16 *)take((size_t)64 * d * 2);
  w.hb16 = (bf16 *)take((size_t)64 * dff * 2);
  w.gemm_ws_bytes = gemm_decode_workspace((int)d, (int)dff);
  w.gemm_ws = take(w.gemm_ws_bytes);
  w.stats = (float *)take((size_t)kLnFoldMaxCopies * kMaxForcedSplits * 64 * 2 * sizeof(float));
  w.total = (size_t)(p - (char *)base) + 256;
  return w;
}
// tensor-core decode path: bf16 storage, up to 64 rows (one UMMA N tile)
bool use_tc_decode(const vb_decoder_desc &D, int B) {
  return D.wdtype == VB_BF16 && B >= 1 && B <= 64 && getenv("VB_DECODE_SIMT") == nullptr;
}
}  // namespace

VB_API size_t vb_ar_step_workspace(const vb_decoder_desc *desc, int B, int cache_cap) {
  return carve_step_ws(*desc, B, cache_cap, nullptr).total;
}

namespace {
struct Pending {  // split-K partials of a projection whose bias/residual the next ln_reduce applies
  const float *part = nullptr;
  const float *bias = nullptr;
  int splits = 0, ldp = 0;
};
bool use_pdl() { return getenv("VB_NO_PDL") == nullptr; }

// final LayerNorm + ar_predict_layer + sampler on the tensor-core path
int tc_head(vb_decoder *dec, const vb_ar_head *head, float *x, vb_ar_state *st, const StepWs &w, const Pending &pend,
            cudaStream_t s) {
  const vb_decoder_desc &D = dec->desc;
  const int d = D.d_model, B = st->B;
  const int ldl = (head->n_vocab + 3) & ~3;
  const bool pdl = use_pdl();
  if (head->fold.wf && pend.part == nullptr && tune("VB_DECODE_FOLD", 1) != 0) {
    // final LayerNorm folded into ar_predict_layer: the projection reads the fp32 rows, the sampler applies the moments
    int sp = 1, ldp = 0, cp = 1;
    VB_TRY(launch_gemm_decode_x(x, B, d, (const bf16 *)head->fold.wf, head->n_vocab, d, 0, (float *)w.gemm_ws,
                                w.gemm_ws_bytes, w.stats, &sp, &ldp, &cp, nullptr, pdl, s));
    const LnFoldStats fs{w.stats, head->fold.c, sp, d, 1e-5f, cp};
    return launch_ar_sample(st->logits, ldl, (const float *)w.gemm_ws, sp, ldp, head, st, d, nullptr,
                            head->greedy ? 0 : 1, pdl, s, &fs);
  }
  VB_TRY(launch_ln_reduce(x, d, B, d, pend.part, pend.splits, pend.ldp, pend.bias, D.final_norm_w, D.final_norm_b,
                          1e-5f, w.xn16, pdl, s));
  int sp = 1, ldp = 0;
  VB_TRY(launch_gemm_decode(w.xn16, B, d, (const bf16 *)head->predict_w, head->n_vocab, d, 0, nullptr, DG_F32,
                            st->logits, nullptr, ldl, nullptr, (float *)w.gemm_ws, w.gemm_ws_bytes, &sp, &ldp, nullptr, pdl,
                            s));
  if (head->greedy)
    VB_TRY(launch_ar_sample(st->logits, ldl, sp > 1 ? (const float *)w.gemm_ws : nullptr, sp, ldp, head, st, d, nullptr,
                            0, pdl, s));
  else if (sp > 1)  // logits only (host-side sampling follows): just reduce the partials
    VB_TRY(launch_ar_sample(st->logits, ldl, (const float *)w.gemm_ws, sp, ldp, head, st, d, nullptr, 1, pdl, s));
  return VB_OK;
}
}  // namespace

VB_API int vb_ar_head_step(vb_decoder_t dec, const vb_ar_head *head, const float *h, vb_ar_state *st,
                           void *workspace, size_t workspace_bytes, vb_stream_t stream) {
  VB_CHECK_ARG(dec && head && h && st, "vb_ar_head_step: null argument");
  const vb_decoder_desc &D = dec->desc;
  cudaStream_t s = (cudaStream_t)stream;
  const int d = D.d_model;
  const int ldl = (head->n_vocab + 3) & ~3;
  if (use_tc_decode(D, st->B)) {
    VB_CHECK_ARG(workspace && workspace_bytes >= vb_ar_step_workspace(&D, st->B, st->cache_cap),
                 "vb_ar_head_step: workspace too small");
    StepWs w = carve_step_ws(D, st->B, st->cache_cap, workspace);
    return tc_head(dec, head, const_cast<float *>(h), st, w, Pending{}, s);
  }
  LnParams ln{D.final_norm_w, D.final_norm_b, nullptr, 1e-5f};
  VB_TRY(launch_gemv(h, d, st->B, head->predict_w, D.wdtype, nullptr, head->n_vocab, d, st->logits, ldl, &ln, 0,
                     nullptr, s));
  if (head->greedy) VB_TRY(launch_ar_sample(st->logits, ldl, nullptr, 0, 0, head, st, d, nullptr, 0, false, s));
  return VB_OK;
}

VB_API int vb_ar_push_tokens(const vb_ar_head *head, vb_ar_state *st, const int64_t *sampled, int d,
                             vb_stream_t stream) {
  VB_CHECK_ARG(head && st && sampled, "vb_ar_push_tokens: null argument");
  const int ldl = (head->n_vocab + 3) & ~3;
  return launch_ar_sample(st->logits, ldl, nullptr, 0, 0, head, st, d, sampled, 0, false, (cudaStream_t)stream);
}

VB_API int vb_ar_decode_step(vb_decoder_t dec, const vb_ar_head *head, vb_ar_state *st, void *workspace,
                             size_t workspace_bytes, vb_stream_t stream) {
  VB_CHECK_ARG(dec && head && st, "vb_ar_decode_step: null argument");
  const vb_decoder_desc &D = dec->desc;
  VB_CHECK_ARG(workspace_bytes >= vb_ar_step_workspace(&D, st->B, st->cache_cap),
               "vb_ar_decode_step: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int d = D.d_model, dff = D.d_ff, B = st->B, dt = D.wdtype, hd = d / D.n_head;
  const size_t ts = elem_size(dt);
  StepWs w = carve_step_ws(D, B, st->cache_cap, workspace);
  float *x = st->x_cur;
  if (use_tc_decode(D, B)) {
    // bf16 tensor-core path: LayerNorm(+pending residual) -> swap-AB split-K tcgen05 projections whose
    // partial sums are consumed by the next kernel in the chain (7 launches per layer, PDL-chained)
    const bool pdl = use_pdl();
    // Tuning of the chain (measured on B200 at d=1024 / d_ff=4096 / B=64, profiles/round1_summary.md): split-K wide
    // enough to fill the SMs is not the optimum for the projections whose partial sums a reduce kernel has to add
    // up again (FFN2: 9 splits beat 18, QKV: 5 beat 6, FFN1: 2 beat 4); 40 % of the KV streams prefetched into L2
    // beat 20 / 60 %.
    // (40 % for the 8-launch chain; the folded chain leaves the projections less time ahead of the attention launch:
    //  30-35 % measured 1-2 % better than 40 %)
    const bool fold_on = dec->fold_qkv && dec->fold_ffn1 && head->fold.wf && tune("VB_DECODE_FOLD", 1) != 0;
    const int pf_env = tune("VB_KV_PREFETCH_PCT", fold_on ? 35 : 40);
    const int pf_pct = B >= 16 ? pf_env : 0;
    const int qkv_env = tune("VB_SPLITS_QKV", 0);
    const int out_splits = tune("VB_SPLITS_OUT", 0);   // 0 = fill the SMs
    const int ffn1_env = tune("VB_SPLITS_FFN1", 0);
    const int ffn2_env = tune("VB_SPLITS_FFN2", 0);
    const int qkv_splits = qkv_env > 0 ? qkv_env : std::max(1, std::min(5, d / 128));
    const int ffn2_splits = ffn2_env > 0 ? ffn2_env : std::max(1, std::min(9, dff / 128));
    // folded chain: no reduce kernel pays for more slabs, and with <= 6 k-blocks per CTA the weight ring never wraps
    const int ffn2_fold = ffn2_env > 0 ? ffn2_env : std::max(1, std::min(11, dff / 128));
    const int ffn1_splits = ffn1_env > 0 ? ffn1_env : std::max(1, std::min(2, d / 128));
    float *P = (float *)w.gemm_ws;
    Pending pend;
    // the four projections of the chain each prefetch a quarter of the first pf_pct % of the KV streams that the
    // NEXT attention launch will read (QKV: this layer's, the other three: the following layer's)
    auto kv_slice = [&](int layer, int quarter) {
      KvPrefetch pf{};
      if (pf_pct <= 0) return pf;
      layer %= D.n_layer;
      pf.kbase = (char *)st->kcache + (size_t)layer * st->cache_layer_stride * ts;
      pf.vbase = (char *)st->vcache + (size_t)layer * st->cache_layer_stride * ts;
      pf.seq_stride_bytes = (int64_t)st->cache_seq_stride * (int64_t)ts;
      pf.B = B; pf.H = D.n_head; pf.cap = st->cache_cap; pf.row_bytes = (int)(hd * ts);
      pf.text_len = st->text_len; pf.prompt_len = st->prompt_len; pf.n_gen = st->n_gen;
      pf.lo_pct = pf_pct * quarter / 4; pf.hi_pct = pf_pct * (quarter + 1) / 4;
      return pf;
    };
    if (fold_on) {
      // Folded chain, 6 launches per layer: the residual stream x is assembled in place by the split-K projections that
      // produce it (red.global.add of every split's tile), the projections that consume it read the fp32 rows and carry
      // the LayerNorm in their weights (vb_ln_fold), the rows' moments travel with the partial sums:
      //   QKV'(x) -> attention (+ moments, KV append) -> out-proj (+= x) -> FFN1'(x) -> ReLU reduce (+ moments) -> FFN2 (+= x)
      const int qkv_f = qkv_env > 0 ? qkv_env : std::max(1, std::min(5, d / 128));
      const int ffn1_f = ffn1_env > 0 ? ffn1_env : std::max(1, std::min(4, d / 128));
      const int out_f = out_splits > 0 ? out_splits : 8;
      for (int l = 0; l < D.n_layer; ++l) {
        const vb_layer_params &L = dec->layers[l];
        const vb_ln_fold &Fq = dec->fold_qkv[l], &Ff = dec->fold_ffn1[l];
        const KvPrefetch pf_qkv = kv_slice(l, 3), pf_out = kv_slice(l + 1, 0), pf_f1 = kv_slice(l + 1, 1),
                         pf_f2 = kv_slice(l + 1, 2);
        void *kc = (char *)st->kcache + (size_t)l * st->cache_layer_stride * ts;
        void *vc = (char *)st->vcache + (size_t)l * st->cache_layer_stride * ts;
        int s1 = 1, ldp1 = 0, cp1 = 1;
        VB_TRY(launch_gemm_decode_x(x, B, d, (const bf16 *)Fq.wf, 3 * d, d, qkv_f, P, w.gemm_ws_bytes, w.stats, &s1, &ldp1,
                                    &cp1, &pf_qkv, pdl, s));
        const LnFoldStats fq{w.stats, Fq.c, s1, d, 1e-5f, cp1};
        VB_TRY(launch_attn_decode(w.q, P, s1, ldp1, Fq.dvec, B, D.n_head, hd, kc, vc, dt, st->cache_seq_stride,
                                  st->cache_cap, st->text_len, st->prompt_len, st->n_gen, st->finished, w.att, w.att16,
                                  w.attn_ws, pdl, s, &fq));
        VB_TRY(launch_gemm_decode(w.att16, B, d, (const bf16 *)L.out_proj_w, d, d, out_f, L.out_proj_b, DG_RESIDUAL, x,
                                  nullptr, d, nullptr, nullptr, 0, nullptr, nullptr, &pf_out, pdl, s, true));
        int sf = 1, ldpf = 0, cpf = 1;
        VB_TRY(launch_gemm_decode_x(x, B, d, (const bf16 *)Ff.wf, dff, d, ffn1_f, P, w.gemm_ws_bytes, w.stats, &sf, &ldpf,
                                    &cpf, &pf_f1, pdl, s));
        const LnFoldStats ff{w.stats, Ff.c, sf, d, 1e-5f, cpf};
        VB_TRY(launch_relu_reduce(P, sf, ldpf, Ff.dvec, B, dff, w.hb16, dff, pdl, s, &ff));
        VB_TRY(launch_gemm_decode(w.hb16, B, dff, (const bf16 *)L.lin2_w, d, dff, ffn2_fold, L.lin2_b, DG_RESIDUAL, x,
                                  nullptr, d, nullptr, nullptr, 0, nullptr, nullptr, &pf_f2, pdl, s, true));
      }
      return tc_head(dec, head, x, st, w, Pending{}, s);
    }
    for (int l = 0; l < D.n_layer; ++l) {
      const vb_layer_params &L = dec->layers[l];
      const KvPrefetch pf_qkv = kv_slice(l, 3), pf_out = kv_slice(l + 1, 0), pf_f1 = kv_slice(l + 1, 1),
                       pf_f2 = kv_slice(l + 1, 2);
      void *kc = (char *)st->kcache + (size_t)l * st->cache_layer_stride * ts;
      void *vc = (char *)st->vcache + (size_t)l * st->cache_layer_stride * ts;
      QkvScatter sc{d, hd, w.q, kc, vc, st->cache_seq_stride, st->cache_cap, st->text_len, st->prompt_len, st->n_gen,
                    st->finished};
      VB_TRY(launch_ln_reduce(x, d, B, d, pend.part, pend.splits, pend.ldp, pend.bias, L.norm1_w, L.norm1_b, 1e-5f,
                              w.xn16, pdl, s));
      int s1 = 1, ldp1 = 0;
      VB_TRY(launch_gemm_decode(w.xn16, B, d, (const bf16 *)L.in_proj_w, 3 * d, d, qkv_splits, L.in_proj_b, DG_QKV, nullptr,
                                nullptr, d, &sc, P, w.gemm_ws_bytes, &s1, &ldp1, &pf_qkv, pdl, s));
      VB_TRY(launch_attn_decode(w.q, s1 > 1 ? P : nullptr, s1, ldp1, L.in_proj_b, B, D.n_head, hd, kc, vc, dt,
                                st->cache_seq_stride, st->cache_cap, st->text_len, st->prompt_len, st->n_gen, st->finished,
                                w.att, w.att16, w.attn_ws, pdl, s));
      int s2 = 1, ldp2 = 0;
      VB_TRY(launch_gemm_decode(w.att16, B, d, (const bf16 *)L.out_proj_w, d, d, out_splits, L.out_proj_b, DG_RESIDUAL, x,
                                nullptr, d, nullptr, P, w.gemm_ws_bytes, &s2, &ldp2, &pf_out, pdl, s));
      VB_TRY(launch_ln_reduce(x, d, B, d, s2 > 1 ? P : nullptr, s2, ldp2, L.out_proj_b, L.norm2_w, L.norm2_b, 1e-5f,
                              w.xn16, pdl, s));
      int sf = 1, ldpf = 0;
      VB_TRY(launch_gemm_decode(w.xn16, B, d, (const bf16 *)L.lin1_w, dff, d, ffn1_splits, L.lin1_b, DG_RELU_BF16, nullptr,
                                w.hb16, dff, nullptr, P, w.gemm_ws_bytes, &sf, &ldpf, &pf_f1, pdl, s));
      if (sf > 1) VB_TRY(launch_relu_reduce(P, sf, ldpf, L.lin1_b, B, dff, w.hb16, dff, pdl, s));
      int s3 = 1, ldp3 = 0;
      VB_TRY(launch_gemm_decode(w.hb16, B, dff, (const bf16 *)L.lin2_w, d, dff, ffn2_splits, L.lin2_b, DG_RESIDUAL, x, nullptr,
                                d, nullptr, P, w.gemm_ws_bytes, &s3, &ldp3, &pf_f2, pdl, s));
      pend = Pending{};
      if (s3 > 1) {
        pend.part = P; pend.bias = L.lin2_b; pend.splits = s3; pend.ldp = ldp3;
      }
    }
    return tc_head(dec, head, x, st, w, pend, s);
  }
  for (int l = 0; l < D.n_layer; ++l) {
    const vb_layer_params &P = dec->layers[l];
    void *kc = (char *)st->kcache + (size_t)l * st->cache_layer_stride * ts;
    void *vc = (char *)st->vcache + (size_t)l * st->cache_layer_stride * ts;
    QkvScatter sc{d, hd, w.q, kc, vc, st->cache_seq_stride, st->cache_cap, st->text_len, st->prompt_len, st->n_gen,
                    st->finished};
    LnParams ln1{P.norm1_w, P.norm1_b, nullptr, 1e-5f};
    VB_TRY(launch_gemv(x, d, B, P.in_proj_w, dt, P.in_proj_b, 3 * d, d, nullptr, 0, &ln1, 3, &sc, s));
    VB_TRY(launch_attn_decode(w.q, nullptr, 0, 0, nullptr, B, D.n_head, hd, kc, vc, dt, st->cache_seq_stride,
                              st->cache_cap, st->text_len, st->prompt_len, st->n_gen, st->finished, w.att, nullptr,
                              w.attn_ws, false, s));
    VB_TRY(launch_gemv(w.att, d, B, P.out_proj_w, dt, P.out_proj_b, d, d, x, d, nullptr, 2, nullptr, s));
    LnParams ln2{P.norm2_w, P.norm2_b, nullptr, 1e-5f};
    VB_TRY(launch_gemv(x, d, B, P.lin1_w, dt, P.lin1_b, dff, d, w.hb, dff, &ln2, 1, nullptr, s));
    VB_TRY(launch_gemv(w.hb, dff, B, P.lin2_w, dt, P.lin2_b, d, dff, x, d, nullptr, 2, nullptr, s));
  }
  return vb_ar_head_step(dec, head, x, st, workspace, workspace_bytes, stream);
}

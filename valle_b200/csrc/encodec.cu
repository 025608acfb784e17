// EnCodec 24 kHz (SEANet encoder/decoder + 2-layer LSTM + 8-stage residual VQ) building blocks, fp32.
//
// The reference reaches this arithmetic through the un-vendored PyPI package `encodec`
// (valle/data/tokenizer.py:219-242: EncodecModel.encodec_model_24khz(), set_target_bandwidth(6.0),
// codec.encode / codec.decode; weight-norm stripped at load, :181-208).  The published architecture is
// restated here (causal SConv1d with reflect padding, ELU pre-activations, residual blocks with 1x1
// shortcut, strided down/up-sampling convs, LSTM with skip, Euclidean-codebook RVQ) and checked
// against transformers' EncodecModel, the only implementation available offline (parity
// "unpinned" w.r.t. the PyPI package).
//
// Layout: activations [B, C, T] fp32, time contiguous (PyTorch conv layout).
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vb {
namespace ec {

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

// ---- causal / asymmetric-padded Conv1d (+pre-ELU, +bias, +residual) -----------------------------
constexpr int CO_T = 32, CI_T = 16, T_T = 64, K_MAX = 16;

__global__ void __launch_bounds__(256)
conv1d_kernel(const float *__restrict__ x, int Cin, int Tin, const float *__restrict__ w,
              const float *__restrict__ bias, int Cout, int K, int stride, int dil, int pad_left, int reflect,
              int pre_elu, const float *__restrict__ residual, float *__restrict__ out, int Tout, int in_w) {
  extern __shared__ float smem[];
  float *ws = smem;                      // [CO_T][CI_T][K]
  float *xs = smem + CO_T * CI_T * K_MAX;  // [CI_T][in_w]
  const int b = blockIdx.z, co0 = blockIdx.y * CO_T, t0 = blockIdx.x * T_T;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const float *xb = x + (int64_t)b * Cin * Tin;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int g0 = t0 * stride - pad_left;  // global time index of xs[.][0]
  for (int ci0 = 0; ci0 < Cin; ci0 += CI_T) {
    for (int idx = tid; idx < CO_T * CI_T * K; idx += 256) {
      const int co = idx / (CI_T * K), rem = idx - co * CI_T * K;
      const int ci = rem / K, k = rem - ci * K;
      float v = 0.f;
      if (co0 + co < Cout && ci0 + ci < Cin) v = w[((int64_t)(co0 + co) * Cin + ci0 + ci) * K + k];
      ws[(co * CI_T + ci) * K_MAX + k] = v;
    }
    for (int idx = tid; idx < CI_T * in_w; idx += 256) {
      const int ci = idx / in_w, i = idx - ci * in_w;
      int g = g0 + i;
      float v = 0.f;
      if (ci0 + ci < Cin) {
        if (reflect) {  // F.pad(mode="reflect") index map (pads are < Tin on this path)
          if (g < 0) g = -g;
          if (g >= Tin) g = 2 * (Tin - 1) - g;
        }
        if (g >= 0 && g < Tin) {
          v = xb[(int64_t)(ci0 + ci) * Tin + g];
          if (pre_elu) v = elu1(v);
        }
      }
      xs[ci * in_w + i] = v;
    }
    __syncthreads();
    const int cin = min(CI_T, Cin - ci0);
    for (int ci = 0; ci < cin; ++ci) {
      const float *xr = xs + ci * in_w + tx * stride;
      for (int k = 0; k < K; ++k) {
        const float xv = xr[k * dil];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(ws[((ty * 8 + i) * CI_T + ci) * K_MAX + k], xv, acc[i]);
      }
    }
    __syncthreads();
  }
  const int t = t0 + tx;
  if (t < Tout) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int co = co0 + ty * 8 + i;
      if (co < Cout) {
        const int64_t o = ((int64_t)b * Cout + co) * Tout + t;
        float v = acc[i] + (bias ? bias[co] : 0.f);
        if (residual) v += residual[o];
        out[o] = v;
      }
    }
  }
}

// ---- causal ConvTranspose1d with K = 2*stride (+pre-ELU), right padding trimmed ------------------
// out[b,co,t] = bias[co] + sum_ci ( in[ci][q] w[ci][co][r] + in[ci][q-1] w[ci][co][r+stride] ),
// q = t / stride, r = t % stride, t in [0, Tin*stride)
__global__ void __launch_bounds__(256)
conv_transpose1d_kernel(const float *__restrict__ x, int Cin, int Tin, const float *__restrict__ w,
                        const float *__restrict__ bias, int Cout, int stride, int pre_elu,
                        float *__restrict__ out) {
  const int Tout = Tin * stride, K = 2 * stride;
  const int b = blockIdx.z, co = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= Tout) return;
  const int q = t / stride, r = t - q * stride;
  const float *xb = x + (int64_t)b * Cin * Tin;
  float acc = 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    const float *wr = w + ((int64_t)ci * Cout + co) * K;
    float a = xb[(int64_t)ci * Tin + q];
    if (pre_elu) a = elu1(a);
    acc = fmaf(a, wr[r], acc);
    if (q > 0) {
      float p = xb[(int64_t)ci * Tin + q - 1];
      if (pre_elu) p = elu1(p);
      acc = fmaf(p, wr[r + stride], acc);
    }
  }
  out[((int64_t)b * Cout + co) * Tout + t] = acc + (bias ? bias[co] : 0.f);
}

// ---- LSTM time step: 128 hidden units x 4 batch rows per CTA --------------------------------------
// gates = xproj[t] (W_ih x + b_ih + b_hh, precomputed) + W_hh h_{t-1}; PyTorch gate order i,f,g,o
__global__ void __launch_bounds__(128)
lstm_step_kernel(const float *__restrict__ xproj_t, const float *__restrict__ whh_t /*[H][4H]*/,
                 const float *__restrict__ h_prev /*[B][H] or null*/, float *__restrict__ c /*[B][H]*/,
                 float *__restrict__ h_out /*[B][H]*/, int B, int H) {
  extern __shared__ float hs[];  // [4][H]
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int b0 = blockIdx.y * 4;
  for (int i = threadIdx.x; i < 4 * H; i += 128) {
    const int bb = i / H, k = i - bb * H;
    hs[i] = (h_prev && b0 + bb < B) ? h_prev[(int64_t)(b0 + bb) * H + k] : 0.f;
  }
  __syncthreads();
  if (j >= H) return;
  float acc[4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) acc[g][bb] = 0.f;
  if (h_prev) {
    for (int k = 0; k < H; ++k) {
      const float *wr = whh_t + (int64_t)k * 4 * H + j;
      const float w0 = wr[0], w1 = wr[H], w2 = wr[2 * H], w3 = wr[3 * H];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const float hv = hs[bb * H + k];
        acc[0][bb] = fmaf(w0, hv, acc[0][bb]);
        acc[1][bb] = fmaf(w1, hv, acc[1][bb]);
        acc[2][bb] = fmaf(w2, hv, acc[2][bb]);
        acc[3][bb] = fmaf(w3, hv, acc[3][bb]);
      }
    }
  }
#pragma unroll
  for (int bb = 0; bb < 4; ++bb) {
    const int b = b0 + bb;
    if (b >= B) continue;
    const float *xp = xproj_t + (int64_t)b * 4 * H;
    const float gi = acc[0][bb] + xp[j], gf = acc[1][bb] + xp[H + j];
    const float gg = acc[2][bb] + xp[2 * H + j], go = acc[3][bb] + xp[3 * H + j];
    const float si = 1.f / (1.f + expf(-gi)), sf = 1.f / (1.f + expf(-gf)), so = 1.f / (1.f + expf(-go));
    const float cprev = h_prev ? c[(int64_t)b * H + j] : 0.f;
    const float cn = sf * cprev + si * tanhf(gg);
    c[(int64_t)b * H + j] = cn;
    h_out[(int64_t)b * H + j] = so * tanhf(cn);
  }
}

// ---- residual vector quantisation: 8 frames per CTA, all stages in one launch ---------------------
// per stage: idx = argmax_j -(|r|^2 - 2 r.e_j + |e_j|^2) (first maximum), r -= e_idx
constexpr int RVQ_ROWS = 8;
__global__ void __launch_bounds__(256)
rvq_encode_kernel(const float *__restrict__ x, int64_t n_rows, int dim, int n_q, int n_codes,
                  const float *__restrict__ cb /*[nq][n_codes][dim]*/, const float *__restrict__ cb_t /*[nq][dim][n_codes]*/,
                  const float *__restrict__ cb_sq /*[nq][n_codes]*/, int64_t *__restrict__ codes, int64_t code_row_stride,
                  int64_t code_q_stride) {
  extern __shared__ float sm[];
  float *rs = sm;                     // [RVQ_ROWS][dim] residuals
  float *xx = rs + RVQ_ROWS * dim;    // [RVQ_ROWS]
  float *bv = xx + RVQ_ROWS;          // [8 warps][RVQ_ROWS] best value
  int *bi = reinterpret_cast<int *>(bv + 8 * RVQ_ROWS);  // [8 warps][RVQ_ROWS]
  int *sel = bi + 8 * RVQ_ROWS;       // [RVQ_ROWS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * RVQ_ROWS;
  for (int i = tid; i < RVQ_ROWS * dim; i += 256) {
    const int r = i / dim, k = i - r * dim;
    rs[i] = (r0 + r < n_rows) ? x[(r0 + r) * dim + k] : 0.f;
  }
  __syncthreads();
  for (int q = 0; q < n_q; ++q) {
    if (tid < RVQ_ROWS) {
      float s = 0.f;
      for (int k = 0; k < dim; ++k) s += rs[tid * dim + k] * rs[tid * dim + k];
      xx[tid] = s;
    }
    __syncthreads();
    float best[RVQ_ROWS];
    int besti[RVQ_ROWS];
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) {
      best[r] = -CUDART_INF_F;
      besti[r] = 0x7fffffff;
    }
    const float *et = cb_t + (int64_t)q * dim * n_codes;
    for (int j = tid; j < n_codes; j += 256) {
      float dot[RVQ_ROWS];
#pragma unroll
      for (int r = 0; r < RVQ_ROWS; ++r) dot[r] = 0.f;
      for (int k = 0; k < dim; ++k) {
        const float e = et[(int64_t)k * n_codes + j];
#pragma unroll
        for (int r = 0; r < RVQ_ROWS; ++r) dot[r] = fmaf(rs[r * dim + k], e, dot[r]);
      }
      const float ee = cb_sq[(int64_t)q * n_codes + j];
#pragma unroll
      for (int r = 0; r < RVQ_ROWS; ++r) {
        const float dist = -(xx[r] - 2.f * dot[r] + ee);
        if (dist > best[r] || (dist == best[r] && j < besti[r])) {
          best[r] = dist;
          besti[r] = j;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) {
      float v = best[r];
      int ix = besti[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, ix, o);
        if (v2 > v || (v2 == v && i2 < ix)) {
          v = v2;
          ix = i2;
        }
      }
      if (lane == 0) {
        bv[warp * RVQ_ROWS + r] = v;
        bi[warp * RVQ_ROWS + r] = ix;
      }
    }
    __syncthreads();
    if (tid < RVQ_ROWS) {
      float v = bv[tid];
      int ix = bi[tid];
      for (int w = 1; w < 8; ++w) {
        const float v2 = bv[w * RVQ_ROWS + tid];
        const int i2 = bi[w * RVQ_ROWS + tid];
        if (v2 > v || (v2 == v && i2 < ix)) {
          v = v2;
          ix = i2;
        }
      }
      sel[tid] = ix;
      if (r0 + tid < n_rows) codes[(r0 + tid) * code_row_stride + q * code_q_stride] = ix;
    }
    __syncthreads();
    for (int i = tid; i < RVQ_ROWS * dim; i += 256) {
      const int r = i / dim, k = i - r * dim;
      rs[i] -= cb[((int64_t)q * n_codes + sel[r]) * dim + k];
    }
    __syncthreads();
  }
}

// ---- generic 3-D permute: out[i_p0][i_p1][i_p2] = in[i0][i1][i2] ---------------------------------
__global__ void permute3_kernel(const float *__restrict__ in, int d0, int d1, int d2, int p0, int p1, int p2,
                                float *__restrict__ out) {
  const int64_t n = (int64_t)d0 * d1 * d2;
  const int dims[3] = {d0, d1, d2};
  const int od1 = dims[p1], od2 = dims[p2];
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    int idx[3];
    const int64_t a = o / ((int64_t)od1 * od2);
    const int64_t rem = o - a * od1 * od2;
    idx[p0] = (int)a;
    idx[p1] = (int)(rem / od2);
    idx[p2] = (int)(rem - (int64_t)idx[p1] * od2);
    out[o] = in[((int64_t)idx[0] * d1 + idx[1]) * d2 + idx[2]];
  }
}

}  // namespace ec
}  // namespace vb

using namespace vb;

VB_API int vb_conv1d(const float *x, int B, int Cin, int Tin, const float *w, const float *bias, int Cout, int K,
                     int stride, int dilation, int pad_left, int pad_right, int reflect, int pre_elu,
                     const float *residual, float *out, int Tout, vb_stream_t stream) {
  VB_CHECK_ARG(K >= 1 && K <= ec::K_MAX && stride >= 1 && stride <= 8, "vb_conv1d: K=%d stride=%d unsupported", K, stride);
  VB_CHECK_ARG(Tout == (Tin + pad_left + pad_right - (K - 1) * dilation - 1) / stride + 1,
               "vb_conv1d: Tout=%d inconsistent with Tin=%d pads=(%d,%d) K=%d stride=%d dil=%d", Tout, Tin, pad_left,
               pad_right, K, stride, dilation);
  VB_CHECK_ARG(!reflect || (pad_left < Tin && pad_right < Tin), "vb_conv1d: reflect pad >= length");
  if (B == 0 || Tout <= 0) return VB_OK;
  const int in_w = (ec::T_T - 1) * stride + (K - 1) * dilation + 1;
  const size_t smem = (size_t)(ec::CO_T * ec::CI_T * ec::K_MAX + ec::CI_T * in_w) * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) VB_CUDA(cudaFuncSetAttribute(ec::conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  dim3 grid((Tout + ec::T_T - 1) / ec::T_T, (Cout + ec::CO_T - 1) / ec::CO_T, B);
  ec::conv1d_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, Cin, Tin, w, bias, Cout, K, stride, dilation, pad_left,
                                                              reflect, pre_elu, residual, out, Tout, in_w);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

VB_API int vb_conv_transpose1d(const float *x, int B, int Cin, int Tin, const float *w, const float *bias, int Cout,
                               int K, int stride, int pre_elu, float *out, vb_stream_t stream) {
  VB_CHECK_ARG(K == 2 * stride, "vb_conv_transpose1d: only K == 2*stride (EnCodec upsampling) is built");
  if (B == 0 || Tin == 0) return VB_OK;
  dim3 grid((Tin * stride + 255) / 256, Cout, B);
  ec::conv_transpose1d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, Cin, Tin, w, bias, Cout, stride, pre_elu, out);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

VB_API int vb_lstm_layer(const float *xproj, const float *whh_t, int T, int B, int H, float *h_seq, float *c_state,
                         vb_stream_t stream) {
  VB_CHECK_ARG(H % 128 == 0 && H <= 2048, "vb_lstm_layer: H=%d must be a multiple of 128", H);
  if (T == 0 || B == 0) return VB_OK;
  const size_t smem = (size_t)4 * H * sizeof(float);
  dim3 grid(H / 128, (B + 3) / 4);
  for (int t = 0; t < T; ++t) {
    const float *hp = t == 0 ? nullptr : h_seq + (int64_t)(t - 1) * B * H;
    ec::lstm_step_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(xproj + (int64_t)t * B * 4 * H, whh_t, hp, c_state,
                                                                   h_seq + (int64_t)t * B * H, B, H);
    VB_LAUNCH_CHECK();
  }
  return VB_OK;
}

VB_API int vb_rvq_encode(const float *x, int64_t n_rows, int dim, int n_q, int n_codes, const float *codebooks,
                         const float *codebooks_t, const float *codebook_sq, int64_t *codes, int64_t code_row_stride,
                         int64_t code_q_stride, vb_stream_t stream) {
  VB_CHECK_ARG(dim <= 512 && n_q >= 1, "vb_rvq_encode: bad dim/n_q");
  if (n_rows == 0) return VB_OK;
  const size_t smem = (size_t)(ec::RVQ_ROWS * dim + ec::RVQ_ROWS + 8 * ec::RVQ_ROWS) * sizeof(float) +
                      (size_t)(8 * ec::RVQ_ROWS + ec::RVQ_ROWS) * sizeof(int);
  const unsigned grid = (unsigned)((n_rows + ec::RVQ_ROWS - 1) / ec::RVQ_ROWS);
  ec::rvq_encode_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, n_rows, dim, n_q, n_codes, codebooks, codebooks_t,
                                                                  codebook_sq, codes, code_row_stride, code_q_stride);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

VB_API int vb_permute3(const float *in, int d0, int d1, int d2, int p0, int p1, int p2, float *out, vb_stream_t stream) {
  VB_CHECK_ARG(((1 << p0) | (1 << p1) | (1 << p2)) == 7, "vb_permute3: not a permutation");
  const int64_t n = (int64_t)d0 * d1 * d2;
  if (n == 0) return VB_OK;
  const unsigned grid = (unsigned)((n + 255) / 256 > 65535 * 8 ? 65535 * 8 : (n + 255) / 256);
  ec::permute3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, d0, d1, d2, p0, p1, p2, out);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

// Consumer kernels of the split-K decode projections (gemm_decode.cu).
//   ln_reduce_kernel: x[b,:] += bias + sum_s partials[s][b][:]   (residual add of
//   valle/modules/transformer.py:297-302, partial sums in fixed order), then
//   out16[b,:] = bf16(LayerNorm(x[b,:]))  (transformer.py:57-74) -- the activation rows of the next
//   tensor-core projection.  One warp per row, the row stays in registers.
#include "common.cuh"
#include "kernels.cuh"

namespace vb {

template <int kVecs>
__global__ void __launch_bounds__(128)
ln_reduce_kernel(float *__restrict__ x, int64_t ldx, int B, int d, const float *__restrict__ partials, int splits,
                 int ldp, const float *__restrict__ bias, const float *__restrict__ gamma,
                 const float *__restrict__ beta, float eps, bf16 *__restrict__ out16) {
  pdl_launch_dependents();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_wait();
  if (b >= B) return;
  float *xr = x + (int64_t)b * ldx;
  float4 v[kVecs];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVecs; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      v[i] = *reinterpret_cast<const float4 *>(xr + c);
      if (partials) {
        const float *p = partials + (int64_t)b * ldp + c;
        float4 a = *reinterpret_cast<const float4 *>(p);
        for (int sidx = 1; sidx < splits; ++sidx) {
          const float4 t = *reinterpret_cast<const float4 *>(p + (int64_t)sidx * 64 * ldp);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        if (bias) {
          const float4 bb = *reinterpret_cast<const float4 *>(bias + c);
          a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
        }
        v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
        *reinterpret_cast<float4 *>(xr + c) = v[i];
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVecs; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float a = v[i].x - mean, bq = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      q += (a * a + bq * bq) + (e * e + f * f);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  bf16 *orow = out16 + (int64_t)b * d;
#pragma unroll
  for (int i = 0; i < kVecs; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float4 g = *reinterpret_cast<const float4 *>(gamma + c);
      const float4 be = *reinterpret_cast<const float4 *>(beta + c);
      __nv_bfloat162 p0 = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y);
      __nv_bfloat162 p1 = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t *>(&p0);
      pk.y = *reinterpret_cast<uint32_t *>(&p1);
      *reinterpret_cast<uint2 *>(orow + c) = pk;
    }
  }
}

int launch_ln_reduce(float *x, int64_t ldx, int B, int d, const float *partials, int splits, int ldp,
                     const float *bias, const float *gamma, const float *beta, float eps, bf16 *out16, bool pdl,
                     cudaStream_t s) {
  VB_CHECK_ARG(d % 4 == 0 && ldx % 4 == 0 && d <= 2048, "ln_reduce: bad d=%d", d);
  const dim3 grid((B + 3) / 4), block(128);
  const int vecs = (d + 127) / 128;
  if (vecs <= 2)
    VB_CUDA(launch_kernel(ln_reduce_kernel<2>, grid, block, 0, s, pdl, x, ldx, B, d, partials, splits, ldp, bias,
                          gamma, beta, eps, out16));
  else if (vecs <= 8)
    VB_CUDA(launch_kernel(ln_reduce_kernel<8>, grid, block, 0, s, pdl, x, ldx, B, d, partials, splits, ldp, bias,
                          gamma, beta, eps, out16));
  else
    VB_CUDA(launch_kernel(ln_reduce_kernel<16>, grid, block, 0, s, pdl, x, ldx, B, d, partials, splits, ldp, bias,
                          gamma, beta, eps, out16));
  count_launch();
  return VB_OK;
}

}  // namespace vb

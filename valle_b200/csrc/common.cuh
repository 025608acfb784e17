// Shared helpers for the sm_100a kernels of libvalle_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/valle_b200.h"

#define VB_API extern "C" __attribute__((visibility("default")))

namespace vb {

// ---- error slot (thread-local, SURVEY 8b: no global mutable state but this) -------------
void set_error(const char *fmt, ...);
extern thread_local int64_t g_launches_tls;
void count_launch();

#define VB_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      vb::set_error(__VA_ARGS__);          \
      return VB_ERR_ARG;                   \
    }                                      \
  } while (0)

#define VB_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      vb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return VB_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define VB_LAUNCH_CHECK()      \
  do {                         \
    vb::count_launch();        \
    VB_CUDA(cudaGetLastError()); \
  } while (0)

#define VB_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != VB_OK) return _s; \
  } while (0)

// ---- dtype helpers ------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void unpack(float (&f)[4]) const {
    f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w;
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // bf16 -> f32 is a 16-bit left shift
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// streaming (read-once) 16-byte global load that does not pollute L1
__device__ __forceinline__ uint4 ldg_stream16(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
template <typename T> __device__ __forceinline__ Vec16<T> load_stream(const T *p);
template <> __device__ __forceinline__ Vec16<float> load_stream<float>(const float *p) {
  Vec16<float> v;
  uint4 r = ldg_stream16(p);
  v.raw = make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z),
                      __uint_as_float(r.w));
  return v;
}
template <> __device__ __forceinline__ Vec16<bf16> load_stream<bf16>(const bf16 *p) {
  Vec16<bf16> v;
  v.raw = ldg_stream16(p);
  return v;
}
template <typename T> __device__ __forceinline__ Vec16<T> load_vec(const T *p);
template <> __device__ __forceinline__ Vec16<float> load_vec<float>(const float *p) {
  Vec16<float> v;
  v.raw = *reinterpret_cast<const float4 *>(p);
  return v;
}
template <> __device__ __forceinline__ Vec16<bf16> load_vec<bf16>(const bf16 *p) {
  Vec16<bf16> v;
  v.raw = *reinterpret_cast<const uint4 *>(p);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- programmatic dependent launch (no-ops unless the launch carries the PDL attribute) -------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         bool pdl, dim3 cluster, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster.x;
    attr[n].val.clusterDim.y = cluster.y;
    attr[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                 bool pdl, Args... args) {
  return launch_kernel_cluster(kern, grid, block, smem, s, pdl, dim3(1, 1, 1), args...);
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace vb

// placeholder until the tcgen05/TMEM GEMM lands (next commit)
#include "common.cuh"
#include "kernels.cuh"
namespace vb {
bool tcgen05_gemm_supported(int64_t, int, int, int64_t, int64_t) { return false; }
int launch_gemm_tcgen05(const bf16 *, int64_t, const bf16 *, const float *, void *, int, int64_t, int64_t, int,
                        int, int, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return VB_ERR_UNSUPPORTED;
}
}  // namespace vb

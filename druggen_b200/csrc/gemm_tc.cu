// tcgen05 contractions (placeholder until the tensor-core kernels land: rejects loudly).
#include "common.cuh"
namespace dg {
int rows_gemm_tc(const float*, const float*, int, const float*, int, const float*, float*, long long, int, int, int, cudaStream_t) {
  return fail("tcgen05 rows_gemm is not built in this library");
}
int gemm_tn_tc(const float*, const float*, float*, long long, int, int, int, cudaStream_t) {
  return fail("tcgen05 gemm_tn is not built in this library");
}
}  // namespace dg

// Either side of the encoder path (last two rows / 8f rank 3): the producer of its inputs and the consumer of its
// outputs, both trivially memory-bound and bit-exact by construction.
//   dg_label2onehot : src/data/utils.py:15-23  out = zeros(labels.shape + [dim]); out.scatter_(-1, labels.unsqueeze(-1), 1.)
//                     labels as int64 (what to_dense_adj gives the reference) or uint8 (1-byte wire format for the host->device copy:
//                     1 B per edge instead of the 20 B of its fp32 one-hot row)
//   dg_argmax_last  : inference.py:197-198     torch.max(t, -1)[1]  -- first maximal index, a NaN wins (first NaN), as ATen on CPU
#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {

template <typename L>
__global__ void onehot_kernel(const L* __restrict__ labels, float* __restrict__ out, long long total, int classes) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / classes;
    const int c = (int)(idx - row * classes);
    out[idx] = (long long)labels[row] == (long long)c ? 1.f : 0.f;     // consecutive threads: consecutive floats; labels via L1
  }
}

__global__ void argmax_last_kernel(const float* __restrict__ x, long long* __restrict__ out, long long rows, int C) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float* p = x + r * C;
    float best = p[0];
    int bi = 0;
    for (int c = 1; c < C; ++c) {
      const float v = p[c];
      // strictly greater keeps the FIRST maximum; a NaN beats every number and the first NaN is kept
      if ((v > best && best == best) || (v != v && best == best)) { best = v; bi = c; }
    }
    out[r] = bi;
  }
}

}  // namespace dg

using namespace dg;

extern "C" int dg_label2onehot(const void* labels, int label_bytes, float* out, long long n, int classes, void* stream) {
  if (n < 0 || classes <= 0) return fail("dg_label2onehot: bad shape n=%lld classes=%d", n, classes);
  if (label_bytes != 1 && label_bytes != 8) return fail("dg_label2onehot: labels are uint8 or int64 (label_bytes=%d)", label_bytes);
  if (n == 0) return 0;
  const long long total = n * classes;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (label_bytes == 1)
    onehot_kernel<unsigned char><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)labels, out, total, classes);
  else
    onehot_kernel<long long><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, out, total, classes);
  return check_launch("dg_label2onehot");
}

extern "C" int dg_argmax_last(const float* x, long long* out, long long rows, int C, void* stream) {
  if (rows < 0 || C <= 0) return fail("dg_argmax_last: bad shape rows=%lld C=%d", rows, C);
  if (rows == 0) return 0;
  long long blocks = (rows + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  argmax_last_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, out, rows, C);
  return check_launch("dg_argmax_last");
}

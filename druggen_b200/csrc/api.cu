// C-ABI dispatch for the dense contractions + library info (include/druggen_b200.h).
#include <cstring>
#include <string>

#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {
int rows_gemm_fp32(const float*, const float*, int, const float*, int, const float*, const float*, float*, long long, int, int, cudaStream_t);
int gemm_tn_fp32(const float*, const float*, float*, float*, long long, int, int, cudaStream_t);
int rows_gemm_tc(const void*, const float*, int, const float*, int, const void*, const float*, void*, long long, int, int, int, int, cudaStream_t);
int gemm_tn_tc(const void*, const void*, float*, float*, long long, int, int, int, int, cudaStream_t);
}  // namespace dg

using namespace dg;

// run-time options: plain process-wide ints, read on the host at launch time and passed to the kernel by value
static int g_opts[DG_OPT_COUNT] = {/*DG_OPT_L2_PREFETCH*/ DG_PF_GEMM_TN | DG_PF_ATTN_FWD, /*DG_OPT_ATTN_BWD*/ 0};
namespace dg {
int opt_get(int key) { return key >= 0 && key < DG_OPT_COUNT ? g_opts[key] : 0; }
}  // namespace dg
static bool g_trace = false;
static std::string g_trace_text;
namespace dg {
bool trace_on() { return g_trace; }
void trace_add(const char* line) {
  g_trace_text += line;
  g_trace_text += '\n';
}
}  // namespace dg
extern "C" int dg_debug_trace(int on) {
  g_trace = on != 0;
  g_trace_text.clear();
  return 0;
}
extern "C" long long dg_debug_trace_read(char* buf, long long cap) {
  const long long n = (long long)g_trace_text.size();
  if (buf != nullptr && cap > 0) {
    const long long m = n < cap - 1 ? n : cap - 1;
    std::memcpy(buf, g_trace_text.data(), (size_t)m);
    buf[m] = 0;
    g_trace_text.clear();
  }
  return n;
}
extern "C" int dg_set_option(int key, int value) {
  if (key < 0 || key >= DG_OPT_COUNT) return fail("dg_set_option: unknown key %d", key);
  g_opts[key] = value;
  return 0;
}
extern "C" int dg_get_option(int key) { return opt_get(key); }

extern "C" int dg_abi_version(void) { return DG_ABI_VERSION; }
extern "C" const char* dg_last_error(void) { return err_buf(); }

extern "C" int dg_has_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10;
}

extern "C" int dg_rows_gemm(const void* a, const float* w, int w_is_nk, const float* bias, int relu,
                            const void* gate, const float* resid, void* out, long long R, int K, int N, int prec,
                            int flags, void* stream) {
  DG_TRACE("dg_rows_gemm", a, (const void*)w, w_is_nk, (const void*)bias, relu, gate, (const void*)resid, (const void*)out, R, K, N, prec, flags);
  if (R <= 0 || K <= 0 || N <= 0) return fail("dg_rows_gemm: bad shape R=%lld K=%d N=%d", R, K, N);
  if (prec == DG_PREC_FP32) {
    if (flags) return fail("dg_rows_gemm: bf16 storage flags need the tensor-core precision");
    return rows_gemm_fp32((const float*)a, w, w_is_nk, bias, relu, (const float*)gate, resid, (float*)out, R, K, N, (cudaStream_t)stream);
  }
  if (prec == DG_PREC_BF16 || prec == DG_PREC_BF16X3)
    return rows_gemm_tc(a, w, w_is_nk, bias, relu, gate, resid, out, R, K, N, prec, flags, (cudaStream_t)stream);
  return fail("dg_rows_gemm: unknown precision %d", prec);
}

extern "C" int dg_gemm_tn(const void* a, const void* b, float* out, float* colsum_a, long long R, int M, int N,
                          int prec, int flags, void* stream) {
  DG_TRACE("dg_gemm_tn", a, b, (const void*)out, (const void*)colsum_a, R, M, N, prec, flags);
  if (R <= 0 || M <= 0 || N <= 0) return fail("dg_gemm_tn: bad shape R=%lld M=%d N=%d", R, M, N);
  if (prec == DG_PREC_FP32) {
    if (flags) return fail("dg_gemm_tn: bf16 storage flags need the tensor-core precision");
    return gemm_tn_fp32((const float*)a, (const float*)b, out, colsum_a, R, M, N, (cudaStream_t)stream);
  }
  if (prec == DG_PREC_BF16 || prec == DG_PREC_BF16X3)
    return gemm_tn_tc(a, b, out, colsum_a, R, M, N, prec, flags, (cudaStream_t)stream);
  return fail("dg_gemm_tn: unknown precision %d", prec);
}

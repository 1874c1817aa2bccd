// Shared helpers for the druggen_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdint>

namespace dg {

// ---- host-side error reporting (thread-local text, C-ABI returns non-zero) -------------------
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return 1;
}
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
  return 0;
}
inline int sm_count() {           // of the CURRENT device (one process may drive several: nn.DataParallel replicas)
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

// ---- run-time options (dg_set_option) --------------------------------------------------------
int opt_get(int key);   // api.cu

// ---- dry-run trace (dg_debug_trace, tests) ----------------------------------------------------
// With the trace on, a kernel entry point records its name and arguments and returns without touching the device: the launch
// programs of the block-level entry points (block.cu) can be listed -- and pinned by a test -- on a box without a GPU.
bool trace_on();                    // api.cu
void trace_add(const char* line);   // api.cu
struct TraceLine {
  char buf[640];
  int n = 0;
  void put(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (n < (int)sizeof buf) n += vsnprintf(buf + n, sizeof buf - n, fmt, ap);
    va_end(ap);
  }
  void add(const void* p) { put(" p:%llx", (unsigned long long)reinterpret_cast<uintptr_t>(p)); }
  void add(long long v) { put(" i:%lld", v); }
  void add(int v) { put(" i:%d", v); }
  void add(float v) { put(" f:%g", (double)v); }
};
template <class... A>
inline int trace_call(const char* name, A... a) {
  TraceLine t;
  t.put("%s", name);
  (t.add(a), ...);
  trace_add(t.buf);
  return 0;
}
#define DG_TRACE(...) \
  if (dg::trace_on()) return dg::trace_call(__VA_ARGS__)

// ---- device helpers --------------------------------------------------------------------------
// L2 prefetch of a contiguous global range by the TMA engine (UBLKPF.L2: no registers, no shared memory, the
// issuing thread does not wait).  `src` 16-byte aligned, `bytes` a multiple of 16.  Used one or two tiles ahead of
// the register-staged loaders so that their loads see L2 latency instead of loaded HBM latency.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, long long bytes) {
  if (bytes > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)bytes) : "memory");
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
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 32 contiguous bytes in ONE request (LDG.E.256, sm_100): the register-staged operand loaders read 8 channels per thread;
// as two LDG.128 every warp instruction touched each 32-byte sector half-used (2x the L1<->L2 sector requests)
__device__ __forceinline__ void ld8(const float* p, float4& lo, float4& hi) {
  asm("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
      : "l"(p));
}
// same, with the L2 evict_last priority: rows that the same kernel reads again a tile or two later (the residual)
__device__ __forceinline__ void ld8_keep(const float* p, float4& lo, float4& hi) {
  asm("ld.global.L2::evict_last.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
      : "l"(p));
}


}  // namespace dg

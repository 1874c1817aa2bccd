// Per-molecule fp32 kernels of the graph attention: edge-modulated scores (layers.py:119-125)
// and softmax over key atoms + value aggregation (layers.py:130-134), each with its backward
// and second-order backward.
//
// Work decomposition: one CTA owns (molecule b, a chunk of query atoms i); thread t owns channel
// t (blockDim.x == D), so every global access of the CTA is a contiguous D*4-byte row of the
// [B,N,N,D] edge tensor (coalesced, 128 B per warp instruction) and everything the attention
// reduces over (key atom j for the softmax / dq, query atom i for dk / dv) is a sequential loop
// in one thread: no cross-thread reduction at all.  Sums over i that cross CTAs (dk, dv) are kept
// per CTA in shared memory [N][D] and flushed with one atomicAdd per (j, channel).
#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {

__global__ void modulate_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                    const float* __restrict__ e, float c, float* __restrict__ out,
                                    long long total4, int N, int D4) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += stride) {
    int c4 = (int)(idx % D4);
    long long row = idx / D4;            // (b*N + i)*N + j
    int j = (int)(row % N);
    long long bi = row / N;              // b*N + i
    long long bj = (bi / N) * N + j;     // b*N + j
    float4 ev = ld4(e + idx * 4), qv = ld4(q + (bi * D4 + c4) * 4), kv = ld4(k + (bj * D4 + c4) * 4);
    st4(out + idx * 4, make_float4(c * qv.x * kv.x * (ev.x * ev.x + ev.x), c * qv.y * kv.y * (ev.y * ev.y + ev.y),
                                   c * qv.z * kv.z * (ev.z * ev.z + ev.z), c * qv.w * kv.w * (ev.w * ev.w + ev.w)));
  }
}

// grid (ichunks, B), block D.  smem: acc[N][D]
__global__ void modulate_bwd_kernel(const float* __restrict__ da, const float* __restrict__ q,
                                    const float* __restrict__ k, const float* __restrict__ e, float c,
                                    float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ de,
                                    int N, int D, int irows) {
  extern __shared__ float acc[];
  int t = threadIdx.x, b = blockIdx.y;
  int i0 = blockIdx.x * irows, i1 = min(N, i0 + irows);
  for (int j = 0; j < N; ++j) acc[j * D + t] = 0.f;
  const float* kb = k + (long long)b * N * D;
  for (int i = i0; i < i1; ++i) {
    float qi = q[((long long)b * N + i) * D + t];
    long long base = (((long long)b * N + i) * N) * D + t;
    float sq = 0.f;
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
      float ev = e[base + (long long)j * D], dv = da[base + (long long)j * D], kj = kb[j * D + t];
      float phi = ev * ev + ev, w = c * dv;
      de[base + (long long)j * D] = w * qi * kj * (2.f * ev + 1.f);
      sq += w * phi * kj;
      acc[j * D + t] += w * phi * qi;
    }
    dq[((long long)b * N + i) * D + t] = sq;       // i is owned by exactly one CTA: plain store
  }
  for (int j = 0; j < N; ++j) atomicAdd(dk + ((long long)b * N + j) * D + t, acc[j * D + t]);
}

__global__ void modulate_bwd_bwd_kernel(const float* __restrict__ uq, const float* __restrict__ uk,
                                        const float* __restrict__ ue, const float* __restrict__ da,
                                        const float* __restrict__ q, const float* __restrict__ k,
                                        const float* __restrict__ e, float c, float* __restrict__ g_da,
                                        float* __restrict__ g_q, float* __restrict__ g_k, float* __restrict__ g_e,
                                        int N, int D, int irows) {
  extern __shared__ float acc[];
  int t = threadIdx.x, b = blockIdx.y;
  int i0 = blockIdx.x * irows, i1 = min(N, i0 + irows);
  for (int j = 0; j < N; ++j) acc[j * D + t] = 0.f;
  const float* kb = k + (long long)b * N * D;
  const float* ukb = uk + (long long)b * N * D;
  for (int i = i0; i < i1; ++i) {
    long long bi = ((long long)b * N + i) * D + t;
    float qi = q[bi], uqi = uq[bi];
    long long base = (((long long)b * N + i) * N) * D + t;
    float sq = 0.f;
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
      long long o = base + (long long)j * D;
      float ev = e[o], dv = da[o], uev = ue[o], kj = kb[j * D + t], ukj = ukb[j * D + t];
      float phi = ev * ev + ev, dphi = 2.f * ev + 1.f, mix = kj * uqi + qi * ukj, w = c * dv;
      g_da[o] = c * (phi * mix + qi * kj * dphi * uev);
      g_e[o] = w * (dphi * mix + 2.f * qi * kj * uev);
      sq += w * (phi * ukj + kj * dphi * uev);
      acc[j * D + t] += w * (phi * uqi + qi * dphi * uev);
    }
    g_q[bi] = sq;
  }
  for (int j = 0; j < N; ++j) atomicAdd(g_k + ((long long)b * N + j) * D + t, acc[j * D + t]);
}

// ---- softmax over j + aggregation -----------------------------------------------------------
__global__ void softmax_agg_fwd_kernel(const float* __restrict__ a, const float* __restrict__ v,
                                       float* __restrict__ g, int N, int D, int irows) {
  int t = threadIdx.x, b = blockIdx.y;
  int i0 = blockIdx.x * irows, i1 = min(N, i0 + irows);
  const float* vb = v + (long long)b * N * D;
  for (int i = i0; i < i1; ++i) {
    long long base = (((long long)b * N + i) * N) * D + t;
    float m = -INFINITY;
#pragma unroll 4
    for (int j = 0; j < N; ++j) m = fmaxf(m, a[base + (long long)j * D]);
    float s = 0.f, sv = 0.f;
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
      float p = __expf(a[base + (long long)j * D] - m);
      s += p;
      sv += p * vb[j * D + t];
    }
    g[((long long)b * N + i) * D + t] = sv / s;
  }
}

__global__ void softmax_agg_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ a,
                                       const float* __restrict__ v, float* __restrict__ da, float* __restrict__ dv,
                                       int accumulate, int N, int D, int irows) {
  extern __shared__ float acc[];
  int t = threadIdx.x, b = blockIdx.y;
  int i0 = blockIdx.x * irows, i1 = min(N, i0 + irows);
  for (int j = 0; j < N; ++j) acc[j * D + t] = 0.f;
  const float* vb = v + (long long)b * N * D;
  for (int i = i0; i < i1; ++i) {
    long long base = (((long long)b * N + i) * N) * D + t;
    float m = -INFINITY;
#pragma unroll 4
    for (int j = 0; j < N; ++j) m = fmaxf(m, a[base + (long long)j * D]);
    float s = 0.f, sv = 0.f;
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
      float p = __expf(a[base + (long long)j * D] - m);
      s += p;
      sv += p * vb[j * D + t];
    }
    float inv = 1.f / s, gi = sv * inv, dgi = dg[((long long)b * N + i) * D + t];
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
      float p = __expf(a[base + (long long)j * D] - m) * inv;
      const float val = p * dgi * (vb[j * D + t] - gi);
      da[base + (long long)j * D] = accumulate ? da[base + (long long)j * D] + val : val;
      acc[j * D + t] += p * dgi;
    }
  }
  for (int j = 0; j < N; ++j) atomicAdd(dv + ((long long)b * N + j) * D + t, acc[j * D + t]);
}

__global__ void softmax_agg_bwd_bwd_kernel(const float* __restrict__ ua, const float* __restrict__ uv,
                                           const float* __restrict__ dg, const float* __restrict__ a,
                                           const float* __restrict__ v, float* __restrict__ g_dg,
                                           float* __restrict__ g_a, float* __restrict__ g_v, int N, int D, int irows) {
  extern __shared__ float acc[];
  int t = threadIdx.x, b = blockIdx.y;
  int i0 = blockIdx.x * irows, i1 = min(N, i0 + irows);
  for (int j = 0; j < N; ++j) acc[j * D + t] = 0.f;
  const float* vb = v + (long long)b * N * D;
  const float* uvb = uv + (long long)b * N * D;
  for (int i = i0; i < i1; ++i) {
    long long base = (((long long)b * N + i) * N) * D + t;
    float m = -INFINITY;
#pragma unroll 4
    for (int j = 0; j < N; ++j) m = fmaxf(m, a[base + (long long)j * D]);
    // s = sum p, sv = sum p v, su = sum p ua, suv = sum p ua v, sw = sum p uv
    float s = 0.f, sv = 0.f, su = 0.f, suv = 0.f, sw = 0.f;
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
      long long o = base + (long long)j * D;
      float p = __expf(a[o] - m), vj = vb[j * D + t], uaj = ua[o];
      s += p; sv += p * vj; su += p * uaj; suv += p * uaj * vj; sw += p * uvb[j * D + t];
    }
    float inv = 1.f / s, gi = sv * inv, mm = su * inv;
    float wbar = (suv - gi * su + sw) * inv;
    long long bi = ((long long)b * N + i) * D + t;
    float dgi = dg[bi];
    g_dg[bi] = wbar;
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
      long long o = base + (long long)j * D;
      float p = __expf(a[o] - m) * inv, vj = vb[j * D + t], uaj = ua[o];
      float w = uaj * (vj - gi) + uvb[j * D + t];
      g_a[o] = dgi * p * (w - wbar - mm * (vj - gi));
      acc[j * D + t] += dgi * p * (uaj - mm);
    }
  }
  for (int j = 0; j < N; ++j) atomicAdd(g_v + ((long long)b * N + j) * D + t, acc[j * D + t]);
}

static int mol_ok(int B, int N, int D, bool smem) {
  if (B <= 0 || N <= 0) return fail("bad shape B=%d N=%d", B, N);
  if (D % 32 || D > 1024 || D <= 0) return fail("channel count D=%d unsupported (need D %% 32 == 0, D <= 1024)", D);
  if (B > 65535) return fail("B=%d exceeds the grid.y limit; split the batch", B);
  if (smem && (size_t)N * D * 4 > 200 * 1024) return fail("N*D too large for the per-CTA accumulator");
  return 0;
}
// rows of query atoms per CTA: enough CTAs to fill the machine ~8x, at most N chunks
static int pick_irows(int B, int N) {
  int want_ctas = sm_count() * 8;
  int chunks = (want_ctas + B - 1) / B;
  if (chunks < 1) chunks = 1;
  if (chunks > N) chunks = N;
  return (N + chunks - 1) / chunks;
}
template <typename Kern>
static int set_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  return 0;
}

// vectorised D == 128 versions (attn_second.cu)
bool attn_second_ok(int B, int N, int D);
int modulate_bwd_v4(const float*, const float*, const float*, const float*, float, float*, float*, float*, int, int, cudaStream_t);
int modulate_bwd_bwd_v4(const float*, const float*, const float*, const float*, const float*, const float*, const float*, float,
                        float*, float*, float*, float*, int, int, cudaStream_t);
int softmax_agg_bwd_v4(const float*, const float*, const float*, float*, float*, int, int, int, cudaStream_t);
int softmax_agg_bwd_bwd_v4(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int, int,
                           cudaStream_t);

}  // namespace dg

using namespace dg;

extern "C" int dg_modulate_fwd(const float* q, const float* k, const float* e, float c, float* out, int B, int N,
                               int D, void* stream) {
  DG_TRACE("dg_modulate_fwd", q, k, e, c, out, B, N, D);
  if (mol_ok(B, N, D, false)) return 1;
  long long total4 = (long long)B * N * N * (D / 4);
  long long blocks = (total4 + 255) / 256, cap = (long long)sm_count() * 16;
  modulate_fwd_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(q, k, e, c, out, total4, N, D / 4);
  return check_launch("dg_modulate_fwd");
}

extern "C" int dg_modulate_bwd(const float* da, const float* q, const float* k, const float* e, float c, float* dq,
                               float* dk, float* de, int B, int N, int D, void* stream) {
  DG_TRACE("dg_modulate_bwd", da, q, k, e, c, dq, dk, de, B, N, D);
  if (mol_ok(B, N, D, true)) return 1;
  if (attn_second_ok(B, N, D)) return modulate_bwd_v4(da, q, k, e, c, dq, dk, de, B, N, (cudaStream_t)stream);
  int irows = pick_irows(B, N);
  size_t smem = (size_t)N * D * 4;
  if (set_smem(modulate_bwd_kernel, smem)) return 1;
  dim3 grid((N + irows - 1) / irows, B);
  modulate_bwd_kernel<<<grid, D, smem, (cudaStream_t)stream>>>(da, q, k, e, c, dq, dk, de, N, D, irows);
  return check_launch("dg_modulate_bwd");
}

extern "C" int dg_modulate_bwd_bwd(const float* uq, const float* uk, const float* ue, const float* da,
                                   const float* q, const float* k, const float* e, float c, float* g_da, float* g_q,
                                   float* g_k, float* g_e, int B, int N, int D, void* stream) {
  DG_TRACE("dg_modulate_bwd_bwd", uq, uk, ue, da, q, k, e, c, g_da, g_q, g_k, g_e, B, N, D);
  if (mol_ok(B, N, D, true)) return 1;
  if (attn_second_ok(B, N, D)) return modulate_bwd_bwd_v4(uq, uk, ue, da, q, k, e, c, g_da, g_q, g_k, g_e, B, N, (cudaStream_t)stream);
  int irows = pick_irows(B, N);
  size_t smem = (size_t)N * D * 4;
  if (set_smem(modulate_bwd_bwd_kernel, smem)) return 1;
  dim3 grid((N + irows - 1) / irows, B);
  modulate_bwd_bwd_kernel<<<grid, D, smem, (cudaStream_t)stream>>>(uq, uk, ue, da, q, k, e, c, g_da, g_q, g_k, g_e, N, D, irows);
  return check_launch("dg_modulate_bwd_bwd");
}

extern "C" int dg_softmax_agg_fwd(const float* a, const float* v, float* g, int B, int N, int D, void* stream) {
  DG_TRACE("dg_softmax_agg_fwd", a, v, g, B, N, D);
  if (mol_ok(B, N, D, false)) return 1;
  int irows = pick_irows(B, N);
  dim3 grid((N + irows - 1) / irows, B);
  softmax_agg_fwd_kernel<<<grid, D, 0, (cudaStream_t)stream>>>(a, v, g, N, D, irows);
  return check_launch("dg_softmax_agg_fwd");
}

extern "C" int dg_softmax_agg_bwd(const float* dg_, const float* a, const float* v, float* da, float* dv, int accumulate,
                                  int B, int N, int D, void* stream) {
  DG_TRACE("dg_softmax_agg_bwd", dg_, a, v, da, dv, accumulate, B, N, D);
  if (mol_ok(B, N, D, true)) return 1;
  if (attn_second_ok(B, N, D)) return softmax_agg_bwd_v4(dg_, a, v, da, dv, accumulate, B, N, (cudaStream_t)stream);
  int irows = pick_irows(B, N);
  size_t smem = (size_t)N * D * 4;
  if (set_smem(softmax_agg_bwd_kernel, smem)) return 1;
  dim3 grid((N + irows - 1) / irows, B);
  softmax_agg_bwd_kernel<<<grid, D, smem, (cudaStream_t)stream>>>(dg_, a, v, da, dv, accumulate, N, D, irows);
  return check_launch("dg_softmax_agg_bwd");
}

extern "C" int dg_softmax_agg_bwd_bwd(const float* ua, const float* uv, const float* dg_, const float* a,
                                      const float* v, float* g_dg, float* g_a, float* g_v, int B, int N, int D,
                                      void* stream) {
  DG_TRACE("dg_softmax_agg_bwd_bwd", ua, uv, dg_, a, v, g_dg, g_a, g_v, B, N, D);
  if (mol_ok(B, N, D, true)) return 1;
  if (attn_second_ok(B, N, D)) return softmax_agg_bwd_bwd_v4(ua, uv, dg_, a, v, g_dg, g_a, g_v, B, N, (cudaStream_t)stream);
  int irows = pick_irows(B, N);
  size_t smem = (size_t)N * D * 4;
  if (set_smem(softmax_agg_bwd_bwd_kernel, smem)) return 1;
  dim3 grid((N + irows - 1) / irows, B);
  softmax_agg_bwd_bwd_kernel<<<grid, D, smem, (cudaStream_t)stream>>>(ua, uv, dg_, a, v, g_dg, g_a, g_v, N, D, irows);
  return check_launch("dg_softmax_agg_bwd_bwd");
}

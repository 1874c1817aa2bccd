// Vectorised (D == 128) versions of the un-fused attention primitives that the twice-differentiable path of the
// gradient penalty runs: modulate_bwd, modulate_bwd_bwd, softmax_agg_bwd, softmax_agg_bwd_bwd.
// Same decomposition as attn_scores.cu: CTA = (molecule, chunk of query atoms), 4 warps, lane = 4 channels
// (one warp instruction = one 512-byte edge row), warp w owns a quarter of the key atoms j; sums over j (per
// query atom) are combined across the four warps through shared memory, sums over i (per key atom) accumulate in
// shared memory exclusive to the owning warp and are flushed with one atomicAdd per element per CTA.
// molecule.cu keeps the generic-D scalar versions (and the forward kernels).
#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {
namespace v4 {

constexpr int D = 128;
constexpr int JU = 4;

#define DG_EACH(OP) OP(x) OP(y) OP(z) OP(w)

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct Ctx {
  int lane, w, b, ch, i0, i1, jlo, jhi;
};
__device__ __forceinline__ Ctx make_ctx(int N, int irows) {
  Ctx c;
  c.lane = threadIdx.x & 31; c.w = threadIdx.x >> 5; c.b = blockIdx.y; c.ch = c.lane * 4;
  c.i0 = blockIdx.x * irows; c.i1 = min(N, c.i0 + irows);
  c.jlo = (c.w * N) / 4; c.jhi = ((c.w + 1) * N) / 4;
  return c;
}
// sum a per-warp float4 partial over the 4 warps; result valid in every thread.  `red` = [4][128] floats.
__device__ __forceinline__ float4 cta_sum4(float4 v, float* red, const Ctx& c) {
  st4(red + c.w * D + c.ch, v);
  __syncthreads();
  float4 r = f4add(f4add(ld4(red + c.ch), ld4(red + D + c.ch)), f4add(ld4(red + 2 * D + c.ch), ld4(red + 3 * D + c.ch)));
  __syncthreads();
  return r;
}
__device__ __forceinline__ void flush_acc(const float* acc, float* dst, int N, const Ctx& c) {
  for (int j = c.jlo; j < c.jhi; ++j) {
    const float4 a = ld4(acc + j * D + c.ch);
    float* p = dst + ((long long)c.b * N + j) * D + c.ch;
    atomicAdd(p, a.x); atomicAdd(p + 1, a.y); atomicAdd(p + 2, a.z); atomicAdd(p + 3, a.w);
  }
}

// the edge rows of query atom i + 2 of up to three edge-sized inputs -> L2 (TMA engine; N contiguous rows each): these kernels keep
// 4 rows per warp in flight and 12-16 warps per SM, i.e. they see the loaded HBM latency on every batch (DG_PF_SECOND)
__device__ __forceinline__ void prefetch_rows_ahead(int pf, const Ctx& c, int i, int N, const float* t0, const float* t1 = nullptr,
                                                    const float* t2 = nullptr) {
  if (pf && threadIdx.x == 0 && i + 2 < c.i1) {
    const long long pb = (((long long)c.b * N + i + 2) * N) * D;
    bulk_prefetch_l2(t0 + pb, (long long)N * D * 4);
    if (t1) bulk_prefetch_l2(t1 + pb, (long long)N * D * 4);
    if (t2) bulk_prefetch_l2(t2 + pb, (long long)N * D * 4);
  }
}

// ---- modulate backward: de = da c q k (2e+1); dq_i = c sum_j da phi k_j; dk_j += c sum_i da phi q_i
__global__ void __launch_bounds__(128, 4)
modulate_bwd_kernel(const float* __restrict__ da, const float* __restrict__ q, const float* __restrict__ k,
                    const float* __restrict__ e, float cc, float* __restrict__ dq, float* __restrict__ dk,
                    float* __restrict__ de, int N, int irows, int pf) {
  extern __shared__ __align__(16) float sm[];
  float* red = sm;              // [4][128]
  float* sdk = sm + 4 * D;      // [N][128]
  const Ctx c = make_ctx(N, irows);
  for (int j = c.jlo; j < c.jhi; ++j) st4(sdk + j * D + c.ch, make_float4(0.f, 0.f, 0.f, 0.f));
  const float* kb = k + (long long)c.b * N * D + c.ch;
  for (int i = c.i0; i < c.i1; ++i) {
    const long long bi = ((long long)c.b * N + i) * D + c.ch;
    const float4 qi = ld4(q + bi);
    const long long base = (((long long)c.b * N + i) * N) * D + c.ch;
    prefetch_rows_ahead(pf, c, i, N, e, da);
    float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = c.jlo; j < c.jhi; j += JU) {
      const int n = min(JU, c.jhi - j);
      float4 ev[JU], dv[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) { ev[u] = ld4(e + base + (long long)(j + u) * D); dv[u] = ld4(da + base + (long long)(j + u) * D); }
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) {
          const float4 kj = ld4(kb + (j + u) * D);
          float4 o, gk;
#define DG_OP(t)                                              \
  {                                                           \
    const float phi = ev[u].t * ev[u].t + ev[u].t, w = cc * dv[u].t; \
    o.t = w * qi.t * kj.t * (2.f * ev[u].t + 1.f);            \
    sq.t = fmaf(w * phi, kj.t, sq.t);                         \
    gk.t = w * phi * qi.t;                                    \
  }
          DG_EACH(DG_OP)
#undef DG_OP
          st4(de + base + (long long)(j + u) * D, o);
          st4(sdk + (j + u) * D + c.ch, f4add(ld4(sdk + (j + u) * D + c.ch), gk));
        }
    }
    const float4 tot = cta_sum4(sq, red, c);
    if (c.w == 0) st4(dq + bi, tot);
  }
  flush_acc(sdk, dk, N, c);
}

// ---- modulate second order
__global__ void __launch_bounds__(128, 3)
modulate_bwd_bwd_kernel(const float* __restrict__ uq, const float* __restrict__ uk, const float* __restrict__ ue,
                        const float* __restrict__ da, const float* __restrict__ q, const float* __restrict__ k,
                        const float* __restrict__ e, float cc, float* __restrict__ g_da, float* __restrict__ g_q,
                        float* __restrict__ g_k, float* __restrict__ g_e, int N, int irows, int pf) {
  extern __shared__ __align__(16) float sm[];
  float* red = sm;
  float* sgk = sm + 4 * D;
  const Ctx c = make_ctx(N, irows);
  for (int j = c.jlo; j < c.jhi; ++j) st4(sgk + j * D + c.ch, make_float4(0.f, 0.f, 0.f, 0.f));
  const float* kb = k + (long long)c.b * N * D + c.ch;
  const float* ukb = uk + (long long)c.b * N * D + c.ch;
  for (int i = c.i0; i < c.i1; ++i) {
    const long long bi = ((long long)c.b * N + i) * D + c.ch;
    const float4 qi = ld4(q + bi), uqi = ld4(uq + bi);
    const long long base = (((long long)c.b * N + i) * N) * D + c.ch;
    prefetch_rows_ahead(pf, c, i, N, e, da, ue);
    float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = c.jlo; j < c.jhi; j += JU) {
      const int n = min(JU, c.jhi - j);
      float4 ev[JU], dv[JU], uv[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) {
          ev[u] = ld4(e + base + (long long)(j + u) * D);
          dv[u] = ld4(da + base + (long long)(j + u) * D);
          uv[u] = ld4(ue + base + (long long)(j + u) * D);
        }
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) {
          const float4 kj = ld4(kb + (j + u) * D), ukj = ld4(ukb + (j + u) * D);
          float4 oa, oe, gk;
#define DG_OP(t)                                                                  \
  {                                                                               \
    const float phi = ev[u].t * ev[u].t + ev[u].t, dphi = 2.f * ev[u].t + 1.f;    \
    const float mix = kj.t * uqi.t + qi.t * ukj.t, w = cc * dv[u].t;              \
    oa.t = cc * (phi * mix + qi.t * kj.t * dphi * uv[u].t);                       \
    oe.t = w * (dphi * mix + 2.f * qi.t * kj.t * uv[u].t);                        \
    sq.t += w * (phi * ukj.t + kj.t * dphi * uv[u].t);                            \
    gk.t = w * (phi * uqi.t + qi.t * dphi * uv[u].t);                             \
  }
          DG_EACH(DG_OP)
#undef DG_OP
          st4(g_da + base + (long long)(j + u) * D, oa);
          st4(g_e + base + (long long)(j + u) * D, oe);
          st4(sgk + (j + u) * D + c.ch, f4add(ld4(sgk + (j + u) * D + c.ch), gk));
        }
    }
    const float4 tot = cta_sum4(sq, red, c);
    if (c.w == 0) st4(g_q + bi, tot);
  }
  flush_acc(sgk, g_k, N, c);
}

// combine the four warps' online-softmax partials of one channel: running max M and NS rescaled running sums
template <int NS>
__device__ __forceinline__ void combine_runs(const float* pm, const float* ps, int stride, float& M, float* S) {
  // pm: [4 warps][128]; ps: [NS][4 warps][128] (stride between sums = `stride` floats)
  M = fmaxf(fmaxf(pm[0], pm[D]), fmaxf(pm[2 * D], pm[3 * D]));
#pragma unroll
  for (int k = 0; k < NS; ++k) S[k] = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float sc = __expf(pm[w * D] - M);
#pragma unroll
    for (int k = 0; k < NS; ++k) S[k] = fmaf(ps[k * stride + w * D], sc, S[k]);
  }
}

// ---- softmax-aggregate backward: da = p dg (v - g) (+ da when accumulate); dv_j += sum_i p dg
__global__ void __launch_bounds__(128, 4)
softmax_agg_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ a, const float* __restrict__ v,
                       float* __restrict__ da, float* __restrict__ dv, int accumulate, int N, int irows, int pf) {
  extern __shared__ __align__(16) float sm[];
  float* red = sm;              // [3][4][128]: m, s, sv
  float* sdv = sm + 12 * D;     // [N][128]
  const Ctx c = make_ctx(N, irows);
  for (int j = c.jlo; j < c.jhi; ++j) st4(sdv + j * D + c.ch, make_float4(0.f, 0.f, 0.f, 0.f));
  const float* vb = v + (long long)c.b * N * D + c.ch;
  for (int i = c.i0; i < c.i1; ++i) {
    const long long bi = ((long long)c.b * N + i) * D + c.ch;
    const long long base = (((long long)c.b * N + i) * N) * D + c.ch;
    prefetch_rows_ahead(pf, c, i, N, a, accumulate ? da : nullptr);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), s = make_float4(0.f, 0.f, 0.f, 0.f), sv = s;
    for (int j = c.jlo; j < c.jhi; j += JU) {
      const int n = min(JU, c.jhi - j);
      float4 av[JU], vv[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) { av[u] = ld4(a + base + (long long)(j + u) * D); vv[u] = ld4(vb + (j + u) * D); }
#define DG_OP(t)                                                             \
  {                                                                          \
    float mx = m.t;                                                          \
    _Pragma("unroll") for (int u = 0; u < JU; ++u) if (u < n) mx = fmaxf(mx, av[u].t); \
    const float sc = __expf(m.t - mx);                                       \
    float ss = s.t * sc, aa = sv.t * sc;                                     \
    _Pragma("unroll") for (int u = 0; u < JU; ++u) if (u < n) {               \
      const float p = __expf(av[u].t - mx);                                  \
      ss += p; aa = fmaf(p, vv[u].t, aa);                                    \
    }                                                                        \
    m.t = mx; s.t = ss; sv.t = aa;                                           \
  }
      DG_EACH(DG_OP)
#undef DG_OP
    }
    st4(red + (0 * 4 + c.w) * D + c.ch, m);
    st4(red + (1 * 4 + c.w) * D + c.ch, s);
    st4(red + (2 * 4 + c.w) * D + c.ch, sv);
    __syncthreads();
    float4 M, inv, g;
#define DG_OP(t, o)                                                     \
  {                                                                     \
    float S[2];                                                         \
    combine_runs<2>(red + c.ch + o, red + 4 * D + c.ch + o, 4 * D, M.t, S); \
    inv.t = 1.f / S[0]; g.t = S[1] * inv.t;                             \
  }
    DG_OP(x, 0) DG_OP(y, 1) DG_OP(z, 2) DG_OP(w, 3)
#undef DG_OP
    __syncthreads();
    const float4 dgi = ld4(dg + bi);
    for (int j = c.jlo; j < c.jhi; j += JU) {
      const int n = min(JU, c.jhi - j);
      float4 av[JU], old[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) {
          av[u] = ld4(a + base + (long long)(j + u) * D);
          old[u] = accumulate ? ld4(da + base + (long long)(j + u) * D) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) {
          const float4 vj = ld4(vb + (j + u) * D);
          float4 o, gv;
#define DG_OP(t)                                              \
  {                                                           \
    const float p = __expf(av[u].t - M.t) * inv.t;            \
    o.t = old[u].t + p * dgi.t * (vj.t - g.t);                \
    gv.t = p * dgi.t;                                         \
  }
          DG_EACH(DG_OP)
#undef DG_OP
          st4(da + base + (long long)(j + u) * D, o);
          st4(sdv + (j + u) * D + c.ch, f4add(ld4(sdv + (j + u) * D + c.ch), gv));
        }
    }
  }
  flush_acc(sdv, dv, N, c);
}

// ---- softmax-aggregate second order:  w_j = ua_j (v_j - g) + uv_j;  Wb = sum p w;  mm = sum p ua
//      g_dg = Wb;  g_a = dg p (w - Wb - mm (v - g));  g_v_j += sum_i dg p (ua - mm)
__global__ void __launch_bounds__(128, 3)
softmax_agg_bwd_bwd_kernel(const float* __restrict__ ua, const float* __restrict__ uv, const float* __restrict__ dg,
                           const float* __restrict__ a, const float* __restrict__ v, float* __restrict__ g_dg,
                           float* __restrict__ g_a, float* __restrict__ g_v, int N, int irows, int pf) {
  extern __shared__ __align__(16) float sm[];
  float* red = sm;              // [6][4][128]: m, s, sv, su, suv, sw
  float* sgv = sm + 24 * D;     // [N][128]
  const Ctx c = make_ctx(N, irows);
  for (int j = c.jlo; j < c.jhi; ++j) st4(sgv + j * D + c.ch, make_float4(0.f, 0.f, 0.f, 0.f));
  const float* vb = v + (long long)c.b * N * D + c.ch;
  const float* uvb = uv + (long long)c.b * N * D + c.ch;
  for (int i = c.i0; i < c.i1; ++i) {
    const long long bi = ((long long)c.b * N + i) * D + c.ch;
    const long long base = (((long long)c.b * N + i) * N) * D + c.ch;
    prefetch_rows_ahead(pf, c, i, N, a, ua);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), s = z4, sv = z4, su = z4, suv = z4, sw = z4;
    for (int j = c.jlo; j < c.jhi; j += JU) {
      const int n = min(JU, c.jhi - j);
      float4 av[JU], uav[JU], vv[JU], uvv[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) {
          av[u] = ld4(a + base + (long long)(j + u) * D);
          uav[u] = ld4(ua + base + (long long)(j + u) * D);
          vv[u] = ld4(vb + (j + u) * D);
          uvv[u] = ld4(uvb + (j + u) * D);
        }
#define DG_OP(t)                                                             \
  {                                                                          \
    float mx = m.t;                                                          \
    _Pragma("unroll") for (int u = 0; u < JU; ++u) if (u < n) mx = fmaxf(mx, av[u].t); \
    const float sc = __expf(m.t - mx);                                       \
    float r0 = s.t * sc, r1 = sv.t * sc, r2 = su.t * sc, r3 = suv.t * sc, r4 = sw.t * sc; \
    _Pragma("unroll") for (int u = 0; u < JU; ++u) if (u < n) {               \
      const float p = __expf(av[u].t - mx);                                  \
      r0 += p; r1 = fmaf(p, vv[u].t, r1); r2 = fmaf(p, uav[u].t, r2);        \
      r3 = fmaf(p * uav[u].t, vv[u].t, r3); r4 = fmaf(p, uvv[u].t, r4);      \
    }                                                                        \
    m.t = mx; s.t = r0; sv.t = r1; su.t = r2; suv.t = r3; sw.t = r4;         \
  }
      DG_EACH(DG_OP)
#undef DG_OP
    }
    st4(red + (0 * 4 + c.w) * D + c.ch, m);
    st4(red + (1 * 4 + c.w) * D + c.ch, s);
    st4(red + (2 * 4 + c.w) * D + c.ch, sv);
    st4(red + (3 * 4 + c.w) * D + c.ch, su);
    st4(red + (4 * 4 + c.w) * D + c.ch, suv);
    st4(red + (5 * 4 + c.w) * D + c.ch, sw);
    __syncthreads();
    float4 M, inv, g, mm, wbar;
#define DG_OP(t, o)                                                          \
  {                                                                          \
    float S[5];                                                              \
    combine_runs<5>(red + c.ch + o, red + 4 * D + c.ch + o, 4 * D, M.t, S);  \
    inv.t = 1.f / S[0]; g.t = S[1] * inv.t; mm.t = S[2] * inv.t;             \
    wbar.t = (S[3] - g.t * S[2] + S[4]) * inv.t;                             \
  }
    DG_OP(x, 0) DG_OP(y, 1) DG_OP(z, 2) DG_OP(w, 3)
#undef DG_OP
    __syncthreads();
    const float4 dgi = ld4(dg + bi);
    if (c.w == 0) st4(g_dg + bi, wbar);
    for (int j = c.jlo; j < c.jhi; j += JU) {
      const int n = min(JU, c.jhi - j);
      float4 av[JU], uav[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) { av[u] = ld4(a + base + (long long)(j + u) * D); uav[u] = ld4(ua + base + (long long)(j + u) * D); }
#pragma unroll
      for (int u = 0; u < JU; ++u)
        if (u < n) {
          const float4 vj = ld4(vb + (j + u) * D), uvj = ld4(uvb + (j + u) * D);
          float4 o, gv;
#define DG_OP(t)                                                        \
  {                                                                     \
    const float p = __expf(av[u].t - M.t) * inv.t;                      \
    const float wj = uav[u].t * (vj.t - g.t) + uvj.t;                   \
    o.t = dgi.t * p * (wj - wbar.t - mm.t * (vj.t - g.t));              \
    gv.t = dgi.t * p * (uav[u].t - mm.t);                               \
  }
          DG_EACH(DG_OP)
#undef DG_OP
          st4(g_a + base + (long long)(j + u) * D, o);
          st4(sgv + (j + u) * D + c.ch, f4add(ld4(sgv + (j + u) * D + c.ch), gv));
        }
    }
  }
  flush_acc(sgv, g_v, N, c);
}

static int irows_for(int B, int N, int ctas_per_sm) {
  int want = sm_count() * ctas_per_sm;
  int chunks = (want + B - 1) / B;
  if (chunks < 1) chunks = 1;
  if (chunks > N) chunks = N;
  return (N + chunks - 1) / chunks;
}
template <typename Kern>
static int smem_attr(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace v4

// true when the vectorised kernels take this shape (else molecule.cu's generic kernels run)
bool attn_second_ok(int B, int N, int Dch) { return Dch == 128 && N >= 4 && B <= 65535 && (size_t)(N + 24) * 128 * 4 <= 200 * 1024; }

int modulate_bwd_v4(const float* da, const float* q, const float* k, const float* e, float c, float* dq, float* dk, float* de,
                    int B, int N, cudaStream_t s) {
  const size_t smem = (size_t)(4 + N) * 128 * 4;
  if (v4::smem_attr(v4::modulate_bwd_kernel, smem)) return 1;
  const int irows = v4::irows_for(B, N, 4);
  dim3 grid((N + irows - 1) / irows, B);
  v4::modulate_bwd_kernel<<<grid, 128, smem, s>>>(da, q, k, e, c, dq, dk, de, N, irows, opt_get(DG_OPT_L2_PREFETCH) & DG_PF_SECOND);
  return check_launch("dg_modulate_bwd");
}
int modulate_bwd_bwd_v4(const float* uq, const float* uk, const float* ue, const float* da, const float* q, const float* k,
                        const float* e, float c, float* g_da, float* g_q, float* g_k, float* g_e, int B, int N, cudaStream_t s) {
  const size_t smem = (size_t)(4 + N) * 128 * 4;
  if (v4::smem_attr(v4::modulate_bwd_bwd_kernel, smem)) return 1;
  const int irows = v4::irows_for(B, N, 3);
  dim3 grid((N + irows - 1) / irows, B);
  v4::modulate_bwd_bwd_kernel<<<grid, 128, smem, s>>>(uq, uk, ue, da, q, k, e, c, g_da, g_q, g_k, g_e, N, irows, opt_get(DG_OPT_L2_PREFETCH) & DG_PF_SECOND);
  return check_launch("dg_modulate_bwd_bwd");
}
int softmax_agg_bwd_v4(const float* dg_, const float* a, const float* v, float* da, float* dv, int accumulate, int B, int N,
                       cudaStream_t s) {
  const size_t smem = (size_t)(12 + N) * 128 * 4;
  if (v4::smem_attr(v4::softmax_agg_bwd_kernel, smem)) return 1;
  const int irows = v4::irows_for(B, N, 4);
  dim3 grid((N + irows - 1) / irows, B);
  v4::softmax_agg_bwd_kernel<<<grid, 128, smem, s>>>(dg_, a, v, da, dv, accumulate, N, irows, opt_get(DG_OPT_L2_PREFETCH) & DG_PF_SECOND);
  return check_launch("dg_softmax_agg_bwd");
}
int softmax_agg_bwd_bwd_v4(const float* ua, const float* uv, const float* dg_, const float* a, const float* v, float* g_dg,
                           float* g_a, float* g_v, int B, int N, cudaStream_t s) {
  const size_t smem = (size_t)(24 + N) * 128 * 4;
  if (v4::smem_attr(v4::softmax_agg_bwd_bwd_kernel, smem)) return 1;
  const int irows = v4::irows_for(B, N, 3);
  dim3 grid((N + irows - 1) / irows, B);
  v4::softmax_agg_bwd_bwd_kernel<<<grid, 128, smem, s>>>(ua, uv, dg_, a, v, g_dg, g_a, g_v, N, irows, opt_get(DG_OPT_L2_PREFETCH) & DG_PF_SECOND);
  return check_launch("dg_softmax_agg_bwd_bwd");
}

}  // namespace dg

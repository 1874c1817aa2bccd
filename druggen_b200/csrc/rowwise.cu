// Row-wise fp32 kernels: residual + LayerNorm (forward, backward, second-order backward),
// ReLU gate, column sums.  One warp owns one row; channels are held in registers as float4s
// (D <= 512, D % 4 == 0), reductions over channels are warp shuffles, reductions over rows
// (affine-parameter gradients) are register partials -> shared memory -> one atomicAdd per
// channel per CTA.  All of these are HBM-bound streaming passes: loads/stores are 16-byte
// vectors, fully coalesced (a warp touches one contiguous 512 B row segment per instruction).
//
// Reference lines: nn.LayerNorm at layers.py:185,189-192 with the residual adds of :187,:188,
// :191,:192 folded in; second-order formulas in DESIGN.md.
#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {

constexpr int kMaxV = 4;          // float4s per lane -> D <= 512 (kernels are templated on V <= kMaxV)
constexpr int kRowWarps = 8;      // warps (rows in flight) per CTA

template <int V>
struct RowVecT {
  float4 v[V];
};

template <int V, typename F>
__device__ __forceinline__ void for_each_chunk_t(int D, int lane, F f) {
#pragma unroll
  for (int t = 0; t < V; ++t) {
    int c = (t * 32 + lane) * 4;
    if (c < D) f(t, c);
  }
}

__device__ __forceinline__ float sum4(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// L2 prefetch of this lane's slice of a future row: the streaming kernels below keep only one row per warp in
// registers, so without it a warp has ~1.5 KB in flight and the SM ~50 KB -- about 4 TB/s at 2 us of loaded latency
__device__ __forceinline__ void prefetch_row(const float* p, long long row, long long R, int D, int lane) {
  if (p != nullptr && row < R && lane * 4 < D) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + row * D + lane * 4));
}

// loads z = a (+ b), returns xhat in-place and rstd
template <int V>
__device__ __forceinline__ float load_normalise(RowVecT<V>& z, const float* a, const float* b, long long row,
                                                int D, int lane, float eps) {
  float s = 0.f;
  for_each_chunk_t<V>(D, lane, [&](int t, int c) {
    float4 x = ld4(a + row * D + c);
    if (b) {
      float4 y = ld4(b + row * D + c);
      x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    }
    z.v[t] = x;
    s += sum4(x);
  });
  float mu = warp_sum(s) / D;
  float q = 0.f;
  for_each_chunk_t<V>(D, lane, [&](int t, int) {
    float4& x = z.v[t];
    x.x -= mu; x.y -= mu; x.z -= mu; x.w -= mu;
    q += dot4(x, x);
  });
  float r = rsqrtf(warp_sum(q) / D + eps);
  for_each_chunk_t<V>(D, lane, [&](int t, int) {
    float4& x = z.v[t];
    x.x *= r; x.y *= r; x.z *= r; x.w *= r;
  });
  return r;
}

template <int V>
__global__ void __launch_bounds__(kRowWarps * 32)
add_ln_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ out, long long R, int D, float eps) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * kRowWarps;
  for (long long row = (long long)blockIdx.x * kRowWarps + warp; row < R; row += stride) {
    prefetch_row(a, row + 2 * stride, R, D, lane);
    prefetch_row(b, row + 2 * stride, R, D, lane);
    RowVecT<V> z;
    load_normalise(z, a, b, row, D, lane, eps);
    for_each_chunk_t<V>(D, lane, [&](int t, int c) {
      float4 g = ld4(gamma + c), be = ld4(beta + c), x = z.v[t];
      st4(out + row * D + c, make_float4(x.x * g.x + be.x, x.y * g.y + be.y, x.z * g.z + be.z, x.w * g.w + be.w));
    });
  }
}

// accumulate per-lane column partials across the CTA and flush with one atomic per channel
template <int V>
__device__ __forceinline__ void flush_columns(const RowVecT<V>& acc, float* dst, int D, int lane, int warp, float* sm) {
  // sm: [kRowWarps][D]
  for_each_chunk_t<V>(D, lane, [&](int t, int c) { st4(sm + warp * D + c, acc.v[t]); });
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) s += sm[w * D + c];
    atomicAdd(dst + c, s);
  }
  __syncthreads();
}

template <int V>
__global__ void __launch_bounds__(kRowWarps * 32)
add_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ b,
                  const float* __restrict__ gamma, float* __restrict__ dz, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, long long R, int D, float eps, int accumulate) {
  extern __shared__ float sm[];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowVecT<V> accg, accb;
  for (int t = 0; t < V; ++t) accg.v[t] = accb.v[t] = make_float4(0, 0, 0, 0);
  const long long stride = (long long)gridDim.x * kRowWarps;
  for (long long row = (long long)blockIdx.x * kRowWarps + warp; row < R; row += stride) {
    prefetch_row(a, row + 2 * stride, R, D, lane);
    prefetch_row(b, row + 2 * stride, R, D, lane);
    prefetch_row(dy, row + 2 * stride, R, D, lane);
    RowVecT<V> xh, gh;
    float r = load_normalise(xh, a, b, row, D, lane, eps);
    float s1 = 0.f, s2 = 0.f;
    for_each_chunk_t<V>(D, lane, [&](int t, int c) {
      float4 d = ld4(dy + row * D + c), g = ld4(gamma + c), x = xh.v[t];
      accg.v[t].x += d.x * x.x; accg.v[t].y += d.y * x.y; accg.v[t].z += d.z * x.z; accg.v[t].w += d.w * x.w;
      accb.v[t].x += d.x; accb.v[t].y += d.y; accb.v[t].z += d.z; accb.v[t].w += d.w;
      float4 h = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      gh.v[t] = h;
      s1 += sum4(h);
      s2 += dot4(h, x);
    });
    float c1 = warp_sum(s1) / D, c2 = warp_sum(s2) / D;
    for_each_chunk_t<V>(D, lane, [&](int t, int c) {
      float4 h = gh.v[t], x = xh.v[t];
      float4 o = make_float4(r * (h.x - c1 - x.x * c2), r * (h.y - c1 - x.y * c2), r * (h.z - c1 - x.z * c2), r * (h.w - c1 - x.w * c2));
      if (accumulate) {                  // dz += : a cotangent that already holds another path's contribution
        const float4 p = ld4(dz + row * D + c);
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      st4(dz + row * D + c, o);
    });
  }
  flush_columns(accg, dgamma, D, lane, warp, sm);
  flush_columns(accb, dbeta, D, lane, warp, sm);
}

// P(w) = w - mean(w) - xh * mean(w * xh), given the two means
__device__ __forceinline__ float4 proj4(float4 w, float4 x, float m1, float m2) {
  return make_float4(w.x - m1 - x.x * m2, w.y - m1 - x.y * m2, w.z - m1 - x.z * m2, w.w - m1 - x.w * m2);
}

template <int V>
__global__ void __launch_bounds__(kRowWarps * 32)
add_ln_bwd_bwd_kernel(const float* __restrict__ u, const float* __restrict__ vg, const float* __restrict__ vb,
                      const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ b,
                      const float* __restrict__ gamma, float* __restrict__ g_dy, float* __restrict__ g_z,
                      float* __restrict__ g_gamma, long long R, int D, float eps) {
  extern __shared__ float sm[];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowVecT<V> accg;
  for (int t = 0; t < V; ++t) accg.v[t] = make_float4(0, 0, 0, 0);
  const long long stride = (long long)gridDim.x * kRowWarps;
  for (long long row = (long long)blockIdx.x * kRowWarps + warp; row < R; row += stride) {
    // (three input rows per warp and 71 registers = 24 warps per SM: without the L2 prefetch the kernel ran at 3.4 TB/s)
    prefetch_row(a, row + 2 * stride, R, D, lane);
    prefetch_row(b, row + 2 * stride, R, D, lane);
    prefetch_row(dy, row + 2 * stride, R, D, lane);
    prefetch_row(u, row + 2 * stride, R, D, lane);
    RowVecT<V> xh, gh, uu;
    float r = load_normalise(xh, a, b, row, D, lane, eps);
    float sg = 0.f, sgx = 0.f, su = 0.f, sux = 0.f, sw = 0.f, swx = 0.f;
    for_each_chunk_t<V>(D, lane, [&](int t, int c) {
      float4 d = ld4(dy + row * D + c), g = ld4(gamma + c), x = xh.v[t], uv = ld4(u + row * D + c);
      float4 h = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      gh.v[t] = h;
      uu.v[t] = uv;
      sg += sum4(h); sgx += dot4(h, x);
      su += sum4(uv); sux += dot4(uv, x);
      if (vg) {
        float4 w = ld4(vg + c);
        w = make_float4(w.x * d.x, w.y * d.y, w.z * d.z, w.w * d.w);
        sw += sum4(w); swx += dot4(w, x);
      }
    });
    float mg = warp_sum(sg) / D, c2 = warp_sum(sgx) / D, mu_ = warp_sum(su) / D, bm = warp_sum(sux) / D;
    float mw = 0.f, mwx = 0.f;
    if (vg) { mw = warp_sum(sw) / D; mwx = warp_sum(swx) / D; }
    // am = mean(u * P(gh))
    float sa = 0.f;
    for_each_chunk_t<V>(D, lane, [&](int t, int) { sa += dot4(uu.v[t], proj4(gh.v[t], xh.v[t], mg, c2)); });
    float am = warp_sum(sa) / D;
    float r2 = r * r;
    for_each_chunk_t<V>(D, lane, [&](int t, int c) {
      float4 x = xh.v[t], g = ld4(gamma + c), d = ld4(dy + row * D + c);
      float4 pu = proj4(uu.v[t], x, mu_, bm), pg = proj4(gh.v[t], x, mg, c2);
      float4 od = make_float4(g.x * r * pu.x, g.y * r * pu.y, g.z * r * pu.z, g.w * r * pu.w);
      float4 oz = make_float4(-r2 * (am * x.x + c2 * pu.x + bm * pg.x), -r2 * (am * x.y + c2 * pu.y + bm * pg.y),
                              -r2 * (am * x.z + c2 * pu.z + bm * pg.z), -r2 * (am * x.w + c2 * pu.w + bm * pg.w));
      if (vg) {
        float4 w = ld4(vg + c);
        od.x += w.x * x.x; od.y += w.y * x.y; od.z += w.z * x.z; od.w += w.w * x.w;
        float4 pw = proj4(make_float4(w.x * d.x, w.y * d.y, w.z * d.z, w.w * d.w), x, mw, mwx);
        oz.x += r * pw.x; oz.y += r * pw.y; oz.z += r * pw.z; oz.w += r * pw.w;
      }
      if (vb) {
        float4 w = ld4(vb + c);
        od.x += w.x; od.y += w.y; od.z += w.z; od.w += w.w;
      }
      st4(g_dy + row * D + c, od);
      st4(g_z + row * D + c, oz);
      accg.v[t].x += d.x * r * pu.x; accg.v[t].y += d.y * r * pu.y;
      accg.v[t].z += d.z * r * pu.z; accg.v[t].w += d.w * r * pu.w;
    });
  }
  flush_columns(accg, g_gamma, D, lane, warp, sm);
}

__global__ void gate_mul_kernel(const float* __restrict__ x, const float* __restrict__ ref, float* __restrict__ out,
                                long long n4, long long n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = ld4(x + i * 4), r = ld4(ref + i * 4);
    st4(out + i * 4, make_float4(r.x > 0.f ? a.x : 0.f, r.y > 0.f ? a.y : 0.f, r.z > 0.f ? a.z : 0.f, r.w > 0.f ? a.w : 0.f));
  }
  // tail (n not a multiple of 4)
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = ref[i] > 0.f ? x[i] : 0.f;
}

// column sums: thread = one column, CTA strides over rows; (blockDim.x = 128, blockDim.y = 4)
__global__ void colsum_kernel(const float* __restrict__ a, float* __restrict__ out, long long R, int N) {
  __shared__ float sm[4][128];
  int col = blockIdx.y * 128 + threadIdx.x;
  float s = 0.f;
  if (col < N)
    for (long long r = (long long)blockIdx.x * 4 + threadIdx.y; r < R; r += (long long)gridDim.x * 4) s += a[r * N + col];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) atomicAdd(out + col, sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x]);
}

#define DG_DISPATCH_V(D, CALL)                                  \
  do {                                                          \
    if ((D) <= 128) { constexpr int V = 1; CALL; }              \
    else if ((D) <= 256) { constexpr int V = 2; CALL; }         \
    else { constexpr int V = 4; CALL; }                         \
  } while (0)

static int row_grid(long long R) {
  long long want = (R + kRowWarps - 1) / kRowWarps;
  long long cap = (long long)sm_count() * 16;  // a few waves of 8-warp CTAs; rows are grid-strided
  return (int)(want < cap ? want : cap);
}
static int ln_ok(long long R, int D) {
  if (R <= 0) return fail("rows must be > 0");
  if (D % 4 || D > 128 * kMaxV || D <= 0) return fail("LayerNorm width D=%d unsupported (need D %% 4 == 0, D <= %d)", D, 128 * kMaxV);
  return 0;
}

}  // namespace dg

using namespace dg;

extern "C" int dg_add_ln_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* out,
                             long long R, int D, float eps, void* stream) {
  DG_TRACE("dg_add_ln_fwd", a, b, gamma, beta, out, R, D, eps);
  if (ln_ok(R, D)) return 1;
  DG_DISPATCH_V(D, (add_ln_fwd_kernel<V><<<row_grid(R), kRowWarps * 32, 0, (cudaStream_t)stream>>>(a, b, gamma, beta, out, R, D, eps)));
  return check_launch("dg_add_ln_fwd");
}

extern "C" int dg_add_ln_bwd(const float* dy, const float* a, const float* b, const float* gamma, float* dz,
                             float* dgamma, float* dbeta, long long R, int D, float eps, int accumulate, void* stream) {
  DG_TRACE("dg_add_ln_bwd", dy, a, b, gamma, dz, dgamma, dbeta, R, D, eps, accumulate);
  if (ln_ok(R, D)) return 1;
  DG_DISPATCH_V(D, (add_ln_bwd_kernel<V><<<row_grid(R), kRowWarps * 32, kRowWarps * D * sizeof(float), (cudaStream_t)stream>>>(
      dy, a, b, gamma, dz, dgamma, dbeta, R, D, eps, accumulate)));
  return check_launch("dg_add_ln_bwd");
}

extern "C" int dg_add_ln_bwd_bwd(const float* u, const float* vg, const float* vb, const float* dy, const float* a,
                                 const float* b, const float* gamma, float* g_dy, float* g_z, float* g_gamma,
                                 long long R, int D, float eps, void* stream) {
  DG_TRACE("dg_add_ln_bwd_bwd", u, vg, vb, dy, a, b, gamma, g_dy, g_z, g_gamma, R, D, eps);
  if (ln_ok(R, D)) return 1;
  DG_DISPATCH_V(D, (add_ln_bwd_bwd_kernel<V><<<row_grid(R), kRowWarps * 32, kRowWarps * D * sizeof(float), (cudaStream_t)stream>>>(
      u, vg, vb, dy, a, b, gamma, g_dy, g_z, g_gamma, R, D, eps)));
  return check_launch("dg_add_ln_bwd_bwd");
}

extern "C" int dg_gate_mul(const float* x, const float* ref, float* out, long long n, void* stream) {
  DG_TRACE("dg_gate_mul", x, ref, out, n);
  if (n <= 0) return fail("n must be > 0");
  long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (blocks < 1) blocks = 1;
  gate_mul_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(x, ref, out, n4, n);
  return check_launch("dg_gate_mul");
}

extern "C" int dg_colsum(const float* a, float* out, long long R, int N, void* stream) {
  DG_TRACE("dg_colsum", a, out, R, N);
  if (R <= 0 || N <= 0) return fail("bad shape");
  long long want = (R + 63) / 64;
  long long cap = (long long)sm_count() * 4;
  dim3 grid((unsigned)(want < cap ? want : cap), (N + 127) / 128), block(128, 4);
  colsum_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a, out, R, N);
  return check_launch("dg_colsum");
}

// Fused per-molecule attention-score kernels (fp32), forward and first-order backward of
//   a_ij = c * q_i * k_j * (e_ij^2 + e_ij)          (layers.py:119-125)
//   g_i  = sum_j softmax_j(a_ij) * v_j              (layers.py:130-134)
// in one pass over the edge tensor each:
//   forward : reads e once, writes a (operand of out_e) and g          -- was modulate_fwd + softmax_agg_fwd
//   backward: reads e and the incoming da (out_e path) once, writes de; dq, dk, dv reduced on chip
//                                                                       -- was softmax_agg_bwd + modulate_bwd
// The scores are recomputed from e, q, k (three FMAs) instead of being re-read; the softmax runs
// online (running max / rescaled sums), so statistics are one sweep and the gradient sweep's second
// visit of e hits L2.
//
// Decomposition (D == 128): CTA = (molecule b, chunk of query atoms i), 4 warps; a lane owns 4
// channels (16-byte loads: a warp instruction moves one whole 512-byte edge row), warp w owns a quarter
// of the key atoms j.  Per query atom the four warps combine their softmax partials through shared
// memory; dk_j / dv_j accumulators are exclusive to the warp that owns j (shared memory, no atomics
// inside the loop) and are flushed to global with one atomicAdd per element per CTA.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {

constexpr int kJU = 4;   // key atoms per load batch: 4 rows x (e, da) x 16 B in flight per lane

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float4 f4s(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// one channel of the online-softmax update with `n` new scores
#define DG_ONLINE(ch)                                                            \
  {                                                                              \
    float mx = m.ch;                                                             \
    _Pragma("unroll") for (int u = 0; u < kJU; ++u) if (u < n) mx = fmaxf(mx, av[u].ch); \
    const float sc = __expf(m.ch - mx);                                          \
    float ss = s.ch * sc, aa = acc.ch * sc;                                      \
    _Pragma("unroll") for (int u = 0; u < kJU; ++u) if (u < n) {                  \
      const float p = __expf(av[u].ch - mx);                                     \
      ss += p;                                                                   \
      aa = fmaf(p, vv[u].ch, aa);                                                \
    }                                                                            \
    m.ch = mx; s.ch = ss; acc.ch = aa;                                           \
  }

// combine the 4 warps' (m, s, acc) partials of one channel
__device__ __forceinline__ void combine4(const float* pm, const float* ps, const float* pa, float& M, float& inv, float& g) {
  M = fmaxf(fmaxf(pm[0], pm[128]), fmaxf(pm[256], pm[384]));
  float S = 0.f, A = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float sc = __expf(pm[w * 128] - M);
    S = fmaf(ps[w * 128], sc, S);
    A = fmaf(pa[w * 128], sc, A);
  }
  inv = 1.f / S;
  g = A * inv;
}

// kMode 0: forward from e, scores written (a_out);  1: backward.  (The forward variants that do not write the scores --
// from e, or from the bf16 scores of dg_attn_edge_fwd -- are attn_fwd_warp_kernel below.)
// (Backward, measured on B200 and not kept: requesting the next batch of (e, da) rows before the current one is reduced costs
// 168 registers / three CTAs per SM and ran 20 % slower; bf16 da_in halves its bytes but the sweep is latency-bound -- 1.5 %.)
template <int kMode>
__global__ void __launch_bounds__(128, 4)
attn_scores_kernel(const float* __restrict__ dg, const float* __restrict__ da_in, const float* __restrict__ q,
                   const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ e, float c,
                   float* __restrict__ a_out, float* __restrict__ g_out, float* __restrict__ de, float* __restrict__ dq,
                   float* __restrict__ dk, float* __restrict__ dv, float* __restrict__ stat_m, float* __restrict__ stat_inv,
                   const float* __restrict__ g_in, int N, int irows, int prefetch, int de_bf16) {
  constexpr int D = 128;
  constexpr bool kBwd = kMode == 1;
  extern __shared__ __align__(16) float sm[];
  // forward : red = [2 parity][3 (m,s,acc)][4 warps][128]   (one barrier per query atom)
  // backward: red = [3 (m,s,acc)][4 warps][128] + [4 warps][128] for the dq partials (three barriers per query atom),
  //           then the dk / dv accumulators [N][128] each -> 54 KB at N = 45, four CTAs per SM
  float* red = sm;
  float* sdk = sm + 16 * D;              // (backward only)
  float* sdv = sdk + N * D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.y;
  const int ch = lane * 4;
  const int i0 = blockIdx.x * irows, i1 = min(N, i0 + irows);
  const int jlo = (w * N) / 4, jhi = ((w + 1) * N) / 4;
  if (kBwd) {
    for (int j = jlo; j < jhi; ++j) {
      st4(sdk + j * D + ch, make_float4(0.f, 0.f, 0.f, 0.f));
      st4(sdv + j * D + ch, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  const float* kb = k + (long long)b * N * D + ch;
  const float* vb = v + (long long)b * N * D + ch;
  for (int i = i0; i < i1; ++i) {
    const long long bi = ((long long)b * N + i) * D + ch;
    const float4 cq = f4s(ld4(q + bi), c);
    const long long base = (((long long)b * N + i) * N) * D + ch;
    if (prefetch && threadIdx.x == 0 && i + 2 < i1) {     // the rows of query atom i+2 -> L2 (TMA engine; contiguous N rows)
      const long long pb = (((long long)b * N + i + 2) * N) * D;
      bulk_prefetch_l2(e + pb, (long long)N * D * 4);
      if (kBwd && da_in) bulk_prefetch_l2(da_in + pb, (long long)N * D * 4);
    }
    float* rd = kBwd ? red : red + ((i - i0) & 1) * 3 * 4 * D;
    float4 M, inv, g;
    const bool have_stats = kBwd && stat_m != nullptr;
    if (have_stats) {                                // statistics saved by the forward kernel: no sweep, no barrier
      M = ld4(stat_m + bi); inv = ld4(stat_inv + bi); g = ld4(g_in + bi);
    } else {
    // ---- sweep 1: scores of my key atoms, online softmax partials
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), s = make_float4(0.f, 0.f, 0.f, 0.f), acc = s;
    for (int j = jlo; j < jhi; j += kJU) {
      const int n = min(kJU, jhi - j);
      float4 ev[kJU], av[kJU], vv[kJU];
      {
#pragma unroll
      for (int u = 0; u < kJU; ++u)
        if (u < n) ev[u] = ld4(e + base + (long long)(j + u) * D);
#pragma unroll
      for (int u = 0; u < kJU; ++u)
        if (u < n) {
          const float4 kj = ld4(kb + (j + u) * D);
          vv[u] = ld4(vb + (j + u) * D);
          av[u] = make_float4(cq.x * kj.x * (ev[u].x * ev[u].x + ev[u].x), cq.y * kj.y * (ev[u].y * ev[u].y + ev[u].y),
                              cq.z * kj.z * (ev[u].z * ev[u].z + ev[u].z), cq.w * kj.w * (ev[u].w * ev[u].w + ev[u].w));
          if (kMode == 0 && a_out != nullptr) st4(a_out + base + (long long)(j + u) * D, av[u]);
        }
      }
      DG_ONLINE(x) DG_ONLINE(y) DG_ONLINE(z) DG_ONLINE(w)
    }
    st4(rd + (0 * 4 + w) * D + ch, m);
    st4(rd + (1 * 4 + w) * D + ch, s);
    st4(rd + (2 * 4 + w) * D + ch, acc);
    __syncthreads();
    combine4(rd + ch + 0, rd + 4 * D + ch + 0, rd + 8 * D + ch + 0, M.x, inv.x, g.x);
    combine4(rd + ch + 1, rd + 4 * D + ch + 1, rd + 8 * D + ch + 1, M.y, inv.y, g.y);
    combine4(rd + ch + 2, rd + 4 * D + ch + 2, rd + 8 * D + ch + 2, M.z, inv.z, g.z);
    combine4(rd + ch + 3, rd + 4 * D + ch + 3, rd + 8 * D + ch + 3, M.w, inv.w, g.w);
    }
    if (!kBwd) {
      if (w == 0) {
        st4(g_out + bi, g);
        if (stat_m != nullptr) { st4(stat_m + bi, M); st4(stat_inv + bi, inv); }
      }
      continue;                                      // next i uses the other parity of `red`
    }
    // ---- sweep 2 (backward): gradients for my key atoms
    const float4 dgi = ld4(dg + bi);
    float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint16_t* da16 = reinterpret_cast<const uint16_t*>(da_in);
    const bool da_is_bf16 = (de_bf16 & 4) != 0;
    float4 ev[kJU], din[kJU];
    auto fetch = [&](int j, float4* pe, float4* pd) {
#pragma unroll
      for (int u = 0; u < kJU; ++u)
        if (j + u < jhi) {
          pe[u] = ld4(e + base + (long long)(j + u) * D);
          if (da_in == nullptr) {
            pd[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (da_is_bf16) {       // the out_e path's gradient stored as bf16 (tensor-core mode): 8 bytes per lane
            const uint2 r = *reinterpret_cast<const uint2*>(da16 + base + (long long)(j + u) * D);
            pd[u] = make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xFFFF0000u), __uint_as_float(r.y << 16),
                                __uint_as_float(r.y & 0xFFFF0000u));
          } else {
            pd[u] = ld4(da_in + base + (long long)(j + u) * D);
          }
        }
    };
    for (int j = jlo; j < jhi; j += kJU) {
      const int n = min(kJU, jhi - j);
      fetch(j, ev, din);
#pragma unroll
      for (int u = 0; u < kJU; ++u)
        if (u < n) {
          const float4 kj = ld4(kb + (j + u) * D), vj = ld4(vb + (j + u) * D);
          float4 o, gk, gv;
#define DG_GRAD(chn)                                                              \
  {                                                                               \
    const float phi = ev[u].chn * ev[u].chn + ev[u].chn;                          \
    float ar = cq.chn * kj.chn * phi;                                             \
    if (de_bf16 & 2) ar = __bfloat162float(__float2bfloat16_rn(ar));              \
    const float p = __expf(ar - M.chn) * inv.chn;                                 \
    const float da = din[u].chn + p * dgi.chn * (vj.chn - g.chn);                 \
    o.chn = da * cq.chn * kj.chn * (2.f * ev[u].chn + 1.f);                       \
    sq.chn = fmaf(da * phi, kj.chn, sq.chn);                                      \
    gk.chn = da * phi * cq.chn;                                                   \
    gv.chn = p * dgi.chn;                                                         \
  }
          DG_GRAD(x) DG_GRAD(y) DG_GRAD(z) DG_GRAD(w)
#undef DG_GRAD
          if (de_bf16 & 1)  // de is only ever a contraction operand (dWe, dy): bf16 storage loses nothing in the tensor-core mode
            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(de) + base + (long long)(j + u) * D) =
                make_uint2(pack2_bf16(o.x, o.y), pack2_bf16(o.z, o.w));
          else {
            if (de_bf16 & 8) {           // de += : the cotangent of E already holds the second-order terms (block_backward_backward)
              const float4 pr = ld4(de + base + (long long)(j + u) * D);
              o.x += pr.x; o.y += pr.y; o.z += pr.z; o.w += pr.w;
            }
            st4(de + base + (long long)(j + u) * D, o);
          }
          float4 ak = ld4(sdk + (j + u) * D + ch), avv = ld4(sdv + (j + u) * D + ch);
          st4(sdk + (j + u) * D + ch, make_float4(ak.x + gk.x, ak.y + gk.y, ak.z + gk.z, ak.w + gk.w));
          st4(sdv + (j + u) * D + ch, make_float4(avv.x + gv.x, avv.y + gv.y, avv.z + gv.z, avv.w + gv.w));
        }
    }
    if (have_stats) {
      // statistics came from the forward: nothing in this query atom needs the other warps -- the four partial sums of
      // dq_i go out as one reduction each (dq zero-filled by the caller) and the warps never meet until the flush (-6 %)
      float* pq = dq + bi;
      atomicAdd(pq, c * sq.x); atomicAdd(pq + 1, c * sq.y); atomicAdd(pq + 2, c * sq.z); atomicAdd(pq + 3, c * sq.w);
      continue;
    }
    // dq_i = c * sum over all key atoms: combine the 4 warps (reuse the m-slot of the other parity)
    float* rq = red + 12 * D;
    st4(rq + w * D + ch, sq);
    __syncthreads();
    if (w == 0) {
      float4 t0 = ld4(rq + ch), t1 = ld4(rq + D + ch), t2 = ld4(rq + 2 * D + ch), t3 = ld4(rq + 3 * D + ch);
      st4(dq + bi, make_float4(c * (t0.x + t1.x + t2.x + t3.x), c * (t0.y + t1.y + t2.y + t3.y), c * (t0.z + t1.z + t2.z + t3.z),
                               c * (t0.w + t1.w + t2.w + t3.w)));
    }
    __syncthreads();                                  // statistics / dq buffers are reused by the next query atom
  }
  if (kBwd) {
    for (int j = jlo; j < jhi; ++j) {
      const float4 ak = ld4(sdk + j * D + ch), avv = ld4(sdv + j * D + ch);
      float* pk = dk + ((long long)b * N + j) * D + ch;
      float* pv = dv + ((long long)b * N + j) * D + ch;
      atomicAdd(pk, ak.x); atomicAdd(pk + 1, ak.y); atomicAdd(pk + 2, ak.z); atomicAdd(pk + 3, ak.w);
      atomicAdd(pv, avv.x); atomicAdd(pv + 1, avv.y); atomicAdd(pv + 2, avv.z); atomicAdd(pv + 3, avv.w);
    }
  }
}


// ---- backward with the forward's statistics, TMA-fed ("ring" kernel) -----------------------------------------------------
// The 4-warp kernel above is latency-bound (2.1 TB/s): a warp has 4 rows in flight and stops loading while it reduces them.
// Here the rows never pass through registers on their way in.  Persistent CTA per SM, 8 warps, no CTA-wide synchronisation
// at all: warp w owns the key atoms j in [w JC, w JC + JC) of every molecule the CTA visits, so
//   * k_j, v_j and the dk_j / dv_j accumulators of its JC key atoms live in REGISTERS for a whole molecule (no shared-memory
//     read-modify-write, no atomics: one plain store per molecule),
//   * per query atom i the warp's rows (b, i, j0 .. j0+JC-1) are CONTIGUOUS in e / da: one cp.async.bulk each lands them in the
//     warp's private ring of `depth` stages (mbarrier complete_tx), issued `depth` query atoms ahead by the warp's own lane 0
//     the moment a stage has been read -- ~150 KB per SM in flight independent of the register budget,
//   * dq_i partials (sum over the warp's key atoms) leave as one vector reduction per lane (red.global.add.v4.f32).
// Flags as in attn_scores_kernel<1> (de_bf16 bits: 1 de stored bf16, 2 scores rounded to bf16, 4 da is bf16, 8 de += ).
// Warps per CTA: 8 with <= 6 key atoms per warp (216 registers; default) or 16 with <= 3.  ncu on the 8-warp form: 148 warp
// instructions per row, issue slots 46 % used with two warps per scheduler (stalls: fixed-latency dependencies, MUFU / LDS
// scoreboards) -- yet the 16-warp form measured SLOWER (0.449 vs 0.387 ms per 1.04 M rows): the per-unit work (node vectors, mbarrier
// wait, refill, dq reduction) is paid per 3 rows instead of per 6.  The kernel is instruction-bound, not HBM-bound (DRAM 38 %).

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// (Measured and not kept: fixing the storage variant at compile time -- no per-row tests, the six rows of a unit one basic block that
// the compiler interleaves, 255 registers -- ran within 1 % of this form.)
template <int JC, int kRingWarps>
__global__ void __launch_bounds__(kRingWarps * 32, 1)
attn_scores_bwd_ring_kernel(const float* __restrict__ dg, const void* __restrict__ da_in, const float* __restrict__ q,
                            const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ e, float c,
                            void* __restrict__ de, float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                            const float* __restrict__ stat_m, const float* __restrict__ stat_inv, const float* __restrict__ g_in,
                            int B, int N, int depth, int da_row_bytes, int flags) {
  constexpr int D = 128;
  extern __shared__ __align__(128) uint8_t ring_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, ch = lane * 4;
  const int j0 = w * JC, rows = min(JC, N - j0);
  if (rows <= 0) return;                                  // (no CTA-wide barrier anywhere below)
  const int acc_row_bytes = (flags & 8) ? 512 : 0;        // de += : the prior de rows ride in the ring too (no load inside the row loop)
  const int stage_bytes = JC * (512 + da_row_bytes + acc_row_bytes);
  uint8_t* ring = ring_raw + (size_t)w * depth * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_raw + (size_t)kRingWarps * depth * stage_bytes) + w * depth;
  if (lane == 0) {
    for (int s = 0; s < depth; ++s) tc::mbar_init(&full[s], 1);
    tc::fence_barrier_init();
  }
  __syncwarp();
  const int nb = blockIdx.x < B ? (B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;      // molecules of this CTA
  const long long total = (long long)nb * N;                                             // (molecule, query atom) units
  const uint32_t tx = (uint32_t)rows * (512 + da_row_bytes + acc_row_bytes);
  // lane 0 issues the units in order, `depth` ahead of the one being reduced: (it_, ii_, is_) = (molecule slot, query atom, stage) of
  // the NEXT unit to issue, advanced incrementally (a 64-bit division per unit costs more than a row of the reduction)
  int it_ = 0, ii_ = 0, is_ = 0;
  auto issue_next = [&]() {
    const long long row = ((long long)(blockIdx.x + it_ * gridDim.x) * N + ii_) * N + j0;
    uint8_t* dst = ring + is_ * stage_bytes;
    tc::mbar_expect_tx(&full[is_], tx);
    tc::bulk_g2s(dst, e + row * D, rows * 512, &full[is_]);
    if (da_row_bytes)
      tc::bulk_g2s(dst + JC * 512, reinterpret_cast<const uint8_t*>(da_in) + row * da_row_bytes, rows * da_row_bytes, &full[is_]);
    if (acc_row_bytes)
      tc::bulk_g2s(dst + JC * (512 + da_row_bytes), reinterpret_cast<const float*>(de) + row * D, rows * 512, &full[is_]);
    if (++ii_ == N) { ii_ = 0; ++it_; }
    if (++is_ == depth) is_ = 0;
  };
  if (lane == 0)
    for (long long n = 0; n < depth && n < total; ++n) issue_next();

  float4 kj[JC], vj[JC], ak[JC], av[JC];
  float4 cq, dgi, M, inv, g;
  auto load_node = [&](long long bi, float4& a0, float4& a1, float4& a2, float4& a3, float4& a4) {
    a0 = ld4(q + bi); a1 = ld4(dg + bi); a2 = ld4(stat_m + bi); a3 = ld4(stat_inv + bi); a4 = ld4(g_in + bi);
  };
  if (total > 0) load_node((long long)blockIdx.x * N * D + ch, cq, dgi, M, inv, g);
  int t = 0, i = 0, s = 0;
  uint32_t par = 0;
  for (long long n = 0; n < total; ++n) {
    const int b = blockIdx.x + t * gridDim.x;
    const long long bi = ((long long)b * N + i) * D + ch;
    if (i == 0) {                                          // new molecule: my key atoms' k, v rows; fresh accumulators
#pragma unroll
      for (int r = 0; r < JC; ++r) {
        const long long o = ((long long)b * N + min(j0 + r, N - 1)) * D + ch;
        kj[r] = ld4(k + o); vj[r] = ld4(v + o);
        ak[r] = make_float4(0.f, 0.f, 0.f, 0.f); av[r] = ak[r];
      }
    }
    // the next unit's node vectors are requested before this unit's rows are waited for
    float4 ncq = cq, ndg = dgi, nM = M, ninv = inv, ng = g;
    if (n + 1 < total) {
      const int ni = i + 1 == N ? 0 : i + 1, nt = i + 1 == N ? t + 1 : t;
      load_node(((long long)(blockIdx.x + nt * gridDim.x) * N + ni) * D + ch, ncq, ndg, nM, ninv, ng);
    }
    const float4 cqs = f4s(cq, c);
    tc::mbar_wait(&full[s], par);
    const uint8_t* st = ring + s * stage_bytes;
    const long long row0 = ((long long)b * N + i) * N + j0;
    float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < JC; ++r)
      if (r < rows) {
        const float4 ev = *reinterpret_cast<const float4*>(st + r * 512 + lane * 16);
        float4 din = make_float4(0.f, 0.f, 0.f, 0.f);
        if (da_row_bytes == 256) {
          const uint2 rr = *reinterpret_cast<const uint2*>(st + JC * 512 + r * 256 + lane * 8);
          din = make_float4(__uint_as_float(rr.x << 16), __uint_as_float(rr.x & 0xFFFF0000u), __uint_as_float(rr.y << 16),
                            __uint_as_float(rr.y & 0xFFFF0000u));
        } else if (da_row_bytes == 512) {
          din = *reinterpret_cast<const float4*>(st + JC * 512 + r * 512 + lane * 16);
        }
        float4 o;
#define DG_RGRAD(chn)                                                             \
  {                                                                               \
    const float phi = ev.chn * ev.chn + ev.chn;                                   \
    float ar = cqs.chn * kj[r].chn * phi;                                         \
    if (flags & 2) ar = __bfloat162float(__float2bfloat16_rn(ar));                \
    const float p = __expf(ar - M.chn) * inv.chn;                                 \
    const float da = din.chn + p * dgi.chn * (vj[r].chn - g.chn);                 \
    o.chn = da * cqs.chn * kj[r].chn * (2.f * ev.chn + 1.f);                      \
    sq.chn = fmaf(da * phi, kj[r].chn, sq.chn);                                   \
    ak[r].chn = fmaf(da * phi, cqs.chn, ak[r].chn);                               \
    av[r].chn = fmaf(p, dgi.chn, av[r].chn);                                      \
  }
        DG_RGRAD(x) DG_RGRAD(y) DG_RGRAD(z) DG_RGRAD(w)
#undef DG_RGRAD
        const long long off = (row0 + r) * D + ch;
        if (flags & 1) {
          *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(de) + off) = make_uint2(pack2_bf16(o.x, o.y), pack2_bf16(o.z, o.w));
        } else {
          float* dp = reinterpret_cast<float*>(de) + off;
          if (flags & 8) {
            const float4 pr = *reinterpret_cast<const float4*>(st + JC * (512 + da_row_bytes) + r * 512 + lane * 16);
            o.x += pr.x; o.y += pr.y; o.z += pr.z; o.w += pr.w;
          }
          st4(dp, o);
        }
      }
    __syncwarp();                                          // every lane has read the stage
    if (lane == 0 && n + depth < total) {
      tc::fence_async_smem();                              // (generic-proxy reads before the async-proxy refill)
      issue_next();
    }
    red_add_v4(dq + bi, c * sq.x, c * sq.y, c * sq.z, c * sq.w);
    if (i + 1 == N) {                                      // molecule done: my key atoms' dk, dv (exclusive rows: plain stores)
#pragma unroll
      for (int r = 0; r < JC; ++r)
        if (r < rows) {
          const long long o = ((long long)b * N + j0 + r) * D + ch;
          st4(dk + o, ak[r]); st4(dv + o, av[r]);
        }
      i = 0; ++t;
    } else {
      ++i;
    }
    if (++s == depth) { s = 0; par ^= 1u; }
    cq = ncq; dgi = ndg; M = nM; inv = ninv; g = ng;
  }
}

// Forward, warp per query atom: one warp owns all N key atoms of its (b, i) -- no shared memory, no barrier; lane = 4
// channels; rows in batches of kFU with the NEXT batch's loads issued before the current batch is reduced (the exp-heavy
// reduction of one batch covers the latency of the next).  kSrc 0: scores recomputed from e (a_out optional); 1: bf16 a16.
constexpr int kFU = 4;
static_assert(kFU == kJU, "DG_ONLINE reduces kJU rows per batch");
template <int kSrc>
__global__ void __launch_bounds__(128, 5)
attn_fwd_warp_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ e,
                     float c, float* __restrict__ a_out, float* __restrict__ g_out, float* __restrict__ stat_m,
                     float* __restrict__ stat_inv, int N, int irows, int prefetch) {
  constexpr int D = 128;
  const uint16_t* a16 = reinterpret_cast<const uint16_t*>(e);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.y;
  const int ch = lane * 4;
  const int i0 = blockIdx.x * irows, i1 = min(N, i0 + irows);
  const float* kb = k + (long long)b * N * D + ch;
  const float* vb = v + (long long)b * N * D + ch;
  for (int i = i0 + w; i < i1; i += 4) {
    const long long bi = ((long long)b * N + i) * D + ch;
    const long long base = (((long long)b * N + i) * N) * D + ch;
    if (prefetch && lane == 0 && i + 4 < i1) {       // my next query atom's rows -> L2
      const long long pb = (((long long)b * N + i + 4) * N) * D;
      if (kSrc == 1) bulk_prefetch_l2(a16 + pb, (long long)N * D * 2);
      else bulk_prefetch_l2(e + pb, (long long)N * D * 4);
    }
    float4 cq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kSrc == 0) cq = f4s(ld4(q + bi), c);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), s = make_float4(0.f, 0.f, 0.f, 0.f), acc = s;
    float4 nx[kFU];                                   // raw rows of the next batch (fp32 e, or bf16 a16 in .x/.y)
    auto fetch = [&](int j) {
#pragma unroll
      for (int u = 0; u < kFU; ++u)
        if (j + u < N) {
          if (kSrc == 1) {
            const uint2 r = *reinterpret_cast<const uint2*>(a16 + base + (long long)(j + u) * D);
            nx[u].x = __uint_as_float(r.x); nx[u].y = __uint_as_float(r.y);
          } else {
            nx[u] = ld4(e + base + (long long)(j + u) * D);
          }
        }
    };
    fetch(0);
    for (int j = 0; j < N; j += kFU) {
      const int n = min(kFU, N - j);
      float4 av[kFU], vv[kFU];
#pragma unroll
      for (int u = 0; u < kFU; ++u)
        if (u < n) {
          vv[u] = ld4(vb + (j + u) * D);
          if (kSrc == 1) {
            const uint32_t lo = __float_as_uint(nx[u].x), hi = __float_as_uint(nx[u].y);
            av[u] = make_float4(__uint_as_float(lo << 16), __uint_as_float(lo & 0xFFFF0000u), __uint_as_float(hi << 16),
                                __uint_as_float(hi & 0xFFFF0000u));
          } else {
            const float4 kj = ld4(kb + (j + u) * D), ev = nx[u];
            av[u] = make_float4(cq.x * kj.x * (ev.x * ev.x + ev.x), cq.y * kj.y * (ev.y * ev.y + ev.y),
                                cq.z * kj.z * (ev.z * ev.z + ev.z), cq.w * kj.w * (ev.w * ev.w + ev.w));
            if (a_out != nullptr) st4(a_out + base + (long long)(j + u) * D, av[u]);
          }
        }
      if (j + kFU < N) fetch(j + kFU);
      DG_ONLINE(x) DG_ONLINE(y) DG_ONLINE(z) DG_ONLINE(w)
    }
    const float4 inv = make_float4(1.f / s.x, 1.f / s.y, 1.f / s.z, 1.f / s.w);
    st4(g_out + bi, make_float4(acc.x * inv.x, acc.y * inv.y, acc.z * inv.z, acc.w * inv.w));
    if (stat_m != nullptr) { st4(stat_m + bi, m); st4(stat_inv + bi, inv); }
  }
}


// (Measured on B200 and not kept: the softmax-aggregate forward from the bf16 scores as a TMA-fed ring kernel like the backward above
// -- one cp.async.bulk per query atom into a per-warp two-stage ring, 8 warps per SM -- ran 0.285 ms per 1.04 M rows against 0.125 ms
// for attn_fwd_warp_kernel<1>: the online softmax is a dependent exp chain, and 8 warps per SM hide less of it than the 20 resident
// warps of the register-staged kernel hide of the loads.)

static int attn_ok(int B, int N, int D) {
  if (B <= 0 || N <= 0) return fail("bad shape B=%d N=%d", B, N);
  if (D != 128) return fail("fused attention-score kernels need D == 128 (got %d)", D);
  if (B > 65535) return fail("B=%d exceeds the grid.y limit; split the batch", B);
  if (N < 4) return fail("fused attention-score kernels need N >= 4 (got %d)", N);
  if ((size_t)(2 * N + 16) * D * 4 > 220 * 1024) return fail("N=%d too large for the per-CTA accumulators", N);
  return 0;
}
static int attn_irows(int B, int N, int ctas_per_sm) {
  int want = sm_count() * ctas_per_sm;
  int chunks = (want + B - 1) / B;
  if (chunks < 1) chunks = 1;
  if (chunks > N) chunks = N;
  return (N + chunks - 1) / chunks;
}

}  // namespace dg

using namespace dg;

extern "C" int dg_attn_scores_fwd(const float* q, const float* k, const float* v, const float* e, float c, float* a,
                                  float* g, float* stat_m, float* stat_inv, int B, int N, int D, void* stream) {
  DG_TRACE("dg_attn_scores_fwd", q, k, v, e, c, a, g, stat_m, stat_inv, B, N, D);
  if (attn_ok(B, N, D)) return 1;
  const size_t smem = (size_t)24 * D * 4;
  const int irows = max(4, attn_irows(B, N, 8));        // one query atom per warp at a time: at least 4 per CTA
  dim3 grid((N + irows - 1) / irows, B);
  if ((stat_m == nullptr) != (stat_inv == nullptr)) return fail("dg_attn_scores_fwd: pass both statistics buffers or neither");
  if (a != nullptr) {       // scores written: the 4-warps-per-query-atom kernel measured faster (0.38 vs 0.43 ms at 1.04 M rows)
    const int ir = attn_irows(B, N, 8);
    dim3 g4((N + ir - 1) / ir, B);
    attn_scores_kernel<0><<<g4, 128, smem, (cudaStream_t)stream>>>(nullptr, nullptr, q, k, v, e, c, a, g, nullptr, nullptr, nullptr, nullptr,
                                                                    stat_m, stat_inv, nullptr, N, ir,
                                                                    opt_get(DG_OPT_L2_PREFETCH) & DG_PF_ATTN_FWD, 0);
  } else {
    attn_fwd_warp_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(q, k, v, e, c, a, g, stat_m, stat_inv, N, irows,
                                                                     opt_get(DG_OPT_L2_PREFETCH) & DG_PF_ATTN_FWD);
  }
  return check_launch("dg_attn_scores_fwd");
}

extern "C" int dg_softmax_agg16_fwd(const void* a_bf16, const float* v, float* g, float* stat_m, float* stat_inv, int B, int N,
                                    int D, void* stream) {
  DG_TRACE("dg_softmax_agg16_fwd", a_bf16, v, g, stat_m, stat_inv, B, N, D);
  if (attn_ok(B, N, D)) return 1;
  const size_t smem = (size_t)24 * D * 4;
  const int irows = max(4, attn_irows(B, N, 8));        // one query atom per warp at a time: at least 4 per CTA
  dim3 grid((N + irows - 1) / irows, B);
  if ((stat_m == nullptr) != (stat_inv == nullptr)) return fail("dg_softmax_agg16_fwd: pass both statistics buffers or neither");
  (void)smem;
  attn_fwd_warp_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(v, v, v, (const float*)a_bf16, 1.f, nullptr, g, stat_m, stat_inv, N, irows,
                                                                   opt_get(DG_OPT_L2_PREFETCH) & DG_PF_ATTN_FWD);
  return check_launch("dg_softmax_agg16_fwd");
}

extern "C" int dg_attn_scores_bwd(const float* dg_, const float* da_in, const float* q, const float* k, const float* v,
                                  const float* e, float c, const float* stat_m, const float* stat_inv, const float* g,
                                  void* de, float* dq, float* dk, float* dv, int B, int N, int D, int de_bf16, void* stream) {
  DG_TRACE("dg_attn_scores_bwd", dg_, da_in, q, k, v, e, c, stat_m, stat_inv, g, de, dq, dk, dv, B, N, D, de_bf16);
  if (attn_ok(B, N, D)) return 1;
  if (stat_m != nullptr && (stat_inv == nullptr || g == nullptr)) return fail("dg_attn_scores_bwd: statistics need stat_inv and g too");
  const int ring_opt = opt_get(DG_OPT_ATTN_BWD);
  if (stat_m != nullptr && N <= 48 && ring_opt != 1) {
    // statistics known and N <= 48: the TMA-fed ring kernel, 8 warps x <= 6 key atoms (default) or 16 warps x <= 3 (option 2)
    const int warps = ring_opt == 2 ? 16 : 8;
    const int jc = (N + warps - 1) / warps;
    const int JC = warps == 8 ? (jc <= 2 ? 2 : jc <= 4 ? 4 : 6) : jc;
    const int da_row = da_in == nullptr ? 0 : ((de_bf16 & 4) ? 256 : 512);
    const int stage = JC * (512 + da_row + ((de_bf16 & 8) ? 512 : 0));
    int depth = (200 * 1024) / (warps * stage);
    if (depth > 8) depth = 8;
    const size_t smem_ring = (size_t)warps * depth * stage + warps * 8 * 8;
    auto kern = warps == 8 ? (JC == 2 ? attn_scores_bwd_ring_kernel<2, 8> : JC == 4 ? attn_scores_bwd_ring_kernel<4, 8> : attn_scores_bwd_ring_kernel<6, 8>)
                           : (JC == 1 ? attn_scores_bwd_ring_kernel<1, 16> : JC == 2 ? attn_scores_bwd_ring_kernel<2, 16> : attn_scores_bwd_ring_kernel<3, 16>);
    cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ring);
    if (er != cudaSuccess) return fail("cudaFuncSetAttribute(attn_scores_bwd_ring): %s", cudaGetErrorString(er));
    const int grid_r = B < sm_count() ? B : sm_count();
    kern<<<grid_r, warps * 32, smem_ring, (cudaStream_t)stream>>>(dg_, da_in, q, k, v, e, c, de, dq, dk, dv, stat_m, stat_inv, g, B, N,
                                                                 depth, da_row, de_bf16);
    return check_launch("dg_attn_scores_bwd(ring)");
  }
  const size_t smem = (size_t)(16 + 2 * N) * D * 4;
  const int irows = attn_irows(B, N, 4);
  dim3 grid((N + irows - 1) / irows, B);
  if (smem > 48 * 1024) {
    cudaError_t er = cudaFuncSetAttribute(attn_scores_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (er != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(er));
  }
  attn_scores_kernel<1><<<grid, 128, smem, (cudaStream_t)stream>>>(dg_, da_in, q, k, v, e, c, nullptr, nullptr, (float*)de, dq, dk, dv,
                                                                    const_cast<float*>(stat_m), const_cast<float*>(stat_inv), g, N, irows,
                                                                    opt_get(DG_OPT_L2_PREFETCH) & DG_PF_ATTN_BWD, de_bf16);
  return check_launch("dg_attn_scores_bwd");
}

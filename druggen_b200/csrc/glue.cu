// The callers and data formats either side of the encoder path, all HBM-bound byte / fp32 streaming work:
//
//   dg_embed_labels_fwd/bwd : models.py:91-94,196-199 for ONE-HOT inputs.  edge_layers / node_layers applied to a one-hot row
//                             is a row of a [classes,128] table (computed from the weights by the host, 5 / 13 rows), so the
//                             prologue + symmetrisation of a label batch is  y_ij = (T[a_ij] + T[a_ji]) / 2  written once from
//                             1-byte labels (instead of two SIMT sgemms + ReLU + permute + add + div over [B,N,N,128]); the backward
//                             is a segmented row sum  dT[l] = sum_ij dy_ij ([a_ij = l] + [a_ji = l]) / 2.
//   dg_gp_interp            : loss.py:21-26  eps * real + (1 - eps) * fake with `real` given as labels (bit-exact with torch's
//                             three elementwise kernels on the one-hot tensor: no contraction into an FMA).
//   dg_gp_penalty / _bwd    : loss.py:42-47  per-sample L2 norm over concat(node, edge gradients), mean((|g| - 1)^2), and its
//                             gradient 2 (|g| - 1) / (B |g|) g.
//   dg_readout_argmax       : models.py:100-101 + inference.py:197-198  logits = x W^T + b over 5 / 13 classes and their argmax
//                             (first maximum) in one pass over the [rows,128] stream; labels leave as int64 or uint8.
//   dg_label2onehot         : src/data/utils.py:15-23 label2onehot from int64 or uint8 (1-byte wire format) labels; a label outside
//                             [0, classes) raises the device flag dg_label_error() reads (torch's scatter_ raises).
//   dg_argmax_last          : inference.py:197-198 torch.max(t, -1)[1], first maximum, a NaN wins -- ATen-CPU's result bit for bit.
//   dg_adamw_flat           : train.py:213-214 torch.optim.AdamW on ONE flat parameter / gradient / moment buffer per network
//                             (decoupled weight decay, bias correction with a per-tensor step count, tensors without a gradient
//                             skipped exactly as torch skips `grad is None`).
#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {

__device__ int g_label_error = 0;      // set by any kernel that meets a label outside [0, classes); read by dg_label_error()

constexpr int kEmbD = 128;             // channel width of the prologue tables (dim)
constexpr int kEmbMaxC = 16;           // label classes: 5 bond types, 13 atom types

template <typename L>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const L* __restrict__ labels, const float* __restrict__ lut,
                                                        float* __restrict__ y, long long rows, int n, int classes, int sym) {
  __shared__ __align__(16) float sT[kEmbMaxC * kEmbD];
  for (int i = threadIdx.x; i < classes * kEmbD; i += blockDim.x) sT[i] = lut[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nn = (long long)n * n;
  for (long long r = warp0; r < rows; r += nwarps) {
    long long l1 = (long long)labels[r], l2 = l1;
    if (sym) {                                                    // r = (b n + i) n + j  ->  (b n + j) n + i
      const long long b = r / nn, ij = r - b * nn;
      const int i = (int)(ij / n), j = (int)(ij - (long long)i * n);
      l2 = (long long)labels[b * nn + (long long)j * n + i];
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l1 < 0 || l1 >= classes || l2 < 0 || l2 >= classes) {
      if (lane == 0) atomicOr(&g_label_error, 1);
    } else if (sym) {
      const float4 a = ld4(sT + l1 * kEmbD + lane * 4), c = ld4(sT + l2 * kEmbD + lane * 4);
      o = make_float4((a.x + c.x) / 2.f, (a.y + c.y) / 2.f, (a.z + c.z) / 2.f, (a.w + c.w) / 2.f);   // models.py:94, same two roundings
    } else {
      o = ld4(sT + l1 * kEmbD + lane * 4);
    }
    st4(y + r * kEmbD + lane * 4, o);
  }
}

// dT[l] += sum over rows with that label; warp-private shared accumulators (lane owns its 4 channels), one atomic flush per block
template <typename L>
__global__ void __launch_bounds__(128) embed_bwd_kernel(const L* __restrict__ labels, const float* __restrict__ dy,
                                                        float* __restrict__ dlut, long long rows, int n, int classes, int sym) {
  __shared__ __align__(16) float sAcc[4][kEmbMaxC * kEmbD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = lane * 4; i < classes * kEmbD; i += 128) st4(&sAcc[warp][i], make_float4(0.f, 0.f, 0.f, 0.f));
  __syncwarp();
  const long long warp0 = (long long)blockIdx.x * 4 + warp, nwarps = (long long)gridDim.x * 4;
  const long long nn = (long long)n * n;
  const float wgt = sym ? 0.5f : 1.f;
  for (long long r0 = warp0 * 4; r0 < rows; r0 += nwarps * 4) {        // 4 rows per trip: their loads are all in flight together
    float4 v[4];
    int l1[4], l2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u;
      l1[u] = l2[u] = -1;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) {
        v[u] = ld4(dy + r * kEmbD + lane * 4);
        long long a = (long long)labels[r], c = a;
        if (sym) {
          const long long b = r / nn, ij = r - b * nn;
          const int i = (int)(ij / n), j = (int)(ij - (long long)i * n);
          c = (long long)labels[b * nn + (long long)j * n + i];
        }
        if (a >= 0 && a < classes && c >= 0 && c < classes) { l1[u] = (int)a; l2[u] = (int)c; }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (l1[u] < 0) continue;
      float* p = &sAcc[warp][l1[u] * kEmbD + lane * 4];
      float4 t = ld4(p);
      t.x += wgt * v[u].x; t.y += wgt * v[u].y; t.z += wgt * v[u].z; t.w += wgt * v[u].w;
      st4(p, t);
      if (sym) {
        p = &sAcc[warp][l2[u] * kEmbD + lane * 4];
        t = ld4(p);
        t.x += wgt * v[u].x; t.y += wgt * v[u].y; t.z += wgt * v[u].z; t.w += wgt * v[u].w;
        st4(p, t);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < classes * kEmbD; i += 128) {
    const float s = (sAcc[0][i] + sAcc[1][i]) + (sAcc[2][i] + sAcc[3][i]);
    if (s != 0.f) atomicAdd(dlut + i, s);
  }
}

// out[row, c] = eps_b * [label == c] + (1 - eps_b) * fake[row, c]; each product / sum rounded as torch's separate kernels round them
template <typename L>
__global__ void gp_interp_kernel(const L* __restrict__ labels, const float* __restrict__ fake, const float* __restrict__ eps,
                                 float* __restrict__ out, long long total, long long rows_per_mol, int classes) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / classes;
    const int c = (int)(idx - row * classes);
    const long long lab = (long long)labels[row];
    if (c == 0 && (lab < 0 || lab >= classes)) atomicOr(&g_label_error, 1);
    const float e = eps[row / rows_per_mol];
    const float real_term = __fmul_rn(e, lab == (long long)c ? 1.f : 0.f);
    out[idx] = __fadd_rn(real_term, __fmul_rn(__fsub_rn(1.f, e), fake[idx]));
  }
}

// per-sample sum of squares over the node and edge gradient rows: one block per molecule
__global__ void __launch_bounds__(256) gp_sqnorm_kernel(const float* __restrict__ g_node, const float* __restrict__ g_edge,
                                                        float* __restrict__ sq, long long len_node, long long len_edge) {
  __shared__ float sred[8];
  const long long b = blockIdx.x;
  float s = 0.f;
  const float* pn = g_node + b * len_node;
  for (long long i = threadIdx.x; i < len_node; i += 256) s = fmaf(pn[i], pn[i], s);
  const float* pe = g_edge + b * len_edge;
  for (long long i = threadIdx.x; i < len_edge; i += 256) s = fmaf(pe[i], pe[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < 8 ? sred[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) sq[b] = t;
  }
}

// penalty = mean_b (sqrt(sq_b) - 1)^2;  coef_b = 2 (|g_b| - 1) / (B |g_b|)   (0 where |g_b| = 0: torch's norm subgradient)
__global__ void __launch_bounds__(256) gp_finish_kernel(const float* __restrict__ sq, float* __restrict__ penalty,
                                                        float* __restrict__ coef, int batch) {
  __shared__ float sred[8];
  float s = 0.f;
  for (int b = threadIdx.x; b < batch; b += 256) {
    const float nrm = sqrtf(sq[b]);
    const float d = nrm - 1.f;
    s = fmaf(d, d, s);
    coef[b] = nrm > 0.f ? 2.f * d / ((float)batch * nrm) : 0.f;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < 8 ? sred[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) penalty[0] = t / (float)batch;
  }
}

// out[b, i] = upstream * coef_b * g[b, i]
__global__ void gp_scale_kernel(const float* __restrict__ g, const float* __restrict__ coef, const float* __restrict__ upstream,
                                float* __restrict__ out, long long total, long long per_mol) {
  const float up = upstream[0];
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = up * coef[idx / per_mol] * g[idx];
}

// logits[r, c] = x[r, :] . w[c, :] + b[c]; idx[r] = first maximal c.  8 lanes per row (16 channels each), 4 rows per warp trip.
template <typename I>
__global__ void __launch_bounds__(256) readout_argmax_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ logits,
                                                             I* __restrict__ idx_out, long long rows, int classes) {
  __shared__ __align__(16) float sW[kEmbMaxC * kEmbD];
  __shared__ float sB[kEmbMaxC];
  for (int i = threadIdx.x; i < classes * kEmbD; i += blockDim.x) sW[i] = w[i];
  if (threadIdx.x < classes) sB[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane >> 3, part = lane & 7;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r0 = warp0 * 4; r0 < rows; r0 += nwarps * 4) {
    const long long r = r0 + sub;
    float4 xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      xv[u] = r < rows ? ld4(x + r * kEmbD + u * 32 + part * 4) : make_float4(0.f, 0.f, 0.f, 0.f);   // 8 lanes x 16 B contiguous per request
    float best = 0.f;
    int bi = 0;
    for (int c = 0; c < classes; ++c) {
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 wv = ld4(sW + c * kEmbD + u * 32 + part * 4);
        s = fmaf(xv[u].x, wv.x, fmaf(xv[u].y, wv.y, fmaf(xv[u].z, wv.z, fmaf(xv[u].w, wv.w, s))));
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += sB[c];
      if (part == 0 && r < rows && logits != nullptr) logits[r * classes + c] = s;
      // strictly greater keeps the FIRST maximum; a NaN beats every number and the first NaN is kept (dg_argmax_last's rule)
      if (c == 0 || (s > best && best == best) || (s != s && best == best)) { best = s; bi = c; }
    }
    if (part == 0 && r < rows && idx_out != nullptr) idx_out[r] = (I)bi;
  }
}

// torch.optim.AdamW (single-tensor formulas, applied per element of a flat buffer); seg: per-tensor [begin, end) and step counts
struct AdamSeg { long long begin, end; float bc1, bc2_sqrt; int active; int pad; };
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, const AdamSeg* __restrict__ segs, int nseg,
                                                         float lr, float beta1, float beta2, float eps, float wd) {
  // one block per (segment, 4096-element chunk) would need a prefix table; segments are few hundred and tiny, so: blockIdx.y = segment
  const AdamSeg s = segs[blockIdx.y];
  if (!s.active) return;
  const float step_size = lr / s.bc1;
  for (long long i = s.begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.end; i += (long long)gridDim.x * blockDim.x) {
    const float grad = g[i];
    float param = p[i];
    param *= 1.f - lr * wd;                                         // decoupled weight decay first (torch _single_tensor_adamw)
    const float mi = m[i] + (grad - m[i]) * (1.f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * grad * grad;    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / s.bc2_sqrt + eps;
    param -= step_size * (mi / denom);
    m[i] = mi; v[i] = vi; p[i] = param;
  }
}

// src/data/utils.py:15-23  out = zeros(labels.shape + [dim]); out.scatter_(-1, labels.unsqueeze(-1), 1.)
template <typename L>
__global__ void onehot_kernel(const L* __restrict__ labels, float* __restrict__ out, long long total, int classes) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / classes;
    const int c = (int)(idx - row * classes);
    const long long lab = (long long)labels[row];
    if (c == 0 && (lab < 0 || lab >= classes)) atomicOr(&g_label_error, 1);   // scatter_ raises: the host turns the flag into an error
    out[idx] = lab == (long long)c ? 1.f : 0.f;     // consecutive threads: consecutive floats; labels via L1
  }
}

// inference.py:197-198  torch.max(t, -1)[1]
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

static int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// models.py:94 for DENSE inputs (the Discriminator on generated / interpolated molecules): out_ij = (e_ij + e_ji) / 2, one pass
// (the reference's permute + add + div are three).  Self-adjoint: its backward is the same launch on the gradient.
__global__ void __launch_bounds__(256) symmetrize_kernel(const float* __restrict__ e, float* __restrict__ out, long long rows, int n, int D) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nn = (long long)n * n;
  for (long long r = warp0; r < rows; r += nwarps) {
    const long long b = r / nn, ij = r - b * nn;
    const int i = (int)(ij / n), j = (int)(ij - (long long)i * n);
    const long long rt = b * nn + (long long)j * n + i;
    for (int c = lane * 4; c < D; c += 128) {
      const float4 a = ld4(e + r * D + c), t = ld4(e + rt * D + c);
      st4(out + r * D + c, make_float4((a.x + t.x) / 2.f, (a.y + t.y) / 2.f, (a.z + t.z) / 2.f, (a.w + t.w) / 2.f));
    }
  }
}

}  // namespace dg

using namespace dg;

extern "C" int dg_label_error(int clear) {
  if (dg::trace_on()) return 0;      // (dry run: nothing was launched, nothing to read back)
  int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_label_error, sizeof(int)) != cudaSuccess) {
    fail("dg_label_error: cannot read the device flag");
    return -1;
  }
  if (v && clear) {
    const int z = 0;
    cudaMemcpyToSymbol(g_label_error, &z, sizeof(int));
  }
  return v;
}

static int embed_check(const char* who, long long rows, int n, int classes, int D, int label_bytes, int sym) {
  if (rows < 0 || n <= 0 || classes <= 0 || classes > kEmbMaxC) return fail("%s: bad shape rows=%lld n=%d classes=%d (classes <= %d)", who, rows, n, classes, kEmbMaxC);
  if (D != kEmbD) return fail("%s: needs D == 128, got %d", who, D);
  if (label_bytes != 1 && label_bytes != 8) return fail("%s: labels are uint8 or int64 (label_bytes=%d)", who, label_bytes);
  if (sym && rows % ((long long)n * n)) return fail("%s: symmetric rows must be a multiple of n*n", who);
  return 0;
}

extern "C" int dg_symmetrize(const float* e, float* out, int B, int N, int D, void* stream) {
  DG_TRACE("dg_symmetrize", e, out, B, N, D);
  if (B <= 0 || N <= 0 || D <= 0 || (D & 3)) return fail("dg_symmetrize: bad shape B=%d N=%d D=%d (D must be a multiple of 4)", B, N, D);
  if (e == out) return fail("dg_symmetrize: in-place is not supported (row ij reads row ji)");
  const long long rows = (long long)B * N * N;
  long long blocks = (rows + 7) / 8, cap = (long long)sm_count() * 16;
  symmetrize_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(e, out, rows, N, D);
  return check_launch("dg_symmetrize");
}

extern "C" int dg_embed_labels_fwd(const void* labels, int label_bytes, const float* lut, float* y, long long rows, int n,
                                   int classes, int D, int sym, void* stream) {
  DG_TRACE("dg_embed_labels_fwd", labels, label_bytes, lut, y, rows, n, classes, D, sym);
  if (embed_check("dg_embed_labels_fwd", rows, n, classes, D, label_bytes, sym)) return 1;
  if (rows == 0) return 0;
  const int grid = grid_for(rows, 8 * 4);
  if (label_bytes == 1)
    embed_fwd_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)labels, lut, y, rows, n, classes, sym);
  else
    embed_fwd_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, lut, y, rows, n, classes, sym);
  return check_launch("dg_embed_labels_fwd");
}

extern "C" int dg_embed_labels_bwd(const void* labels, int label_bytes, const float* dy, float* dlut, long long rows, int n,
                                   int classes, int D, int sym, void* stream) {
  DG_TRACE("dg_embed_labels_bwd", labels, label_bytes, dy, dlut, rows, n, classes, D, sym);
  if (embed_check("dg_embed_labels_bwd", rows, n, classes, D, label_bytes, sym)) return 1;
  if (rows == 0) return 0;
  int grid = grid_for(rows, 4 * 4 * 8);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  if (label_bytes == 1)
    embed_bwd_kernel<unsigned char><<<grid, 128, 0, (cudaStream_t)stream>>>((const unsigned char*)labels, dy, dlut, rows, n, classes, sym);
  else
    embed_bwd_kernel<long long><<<grid, 128, 0, (cudaStream_t)stream>>>((const long long*)labels, dy, dlut, rows, n, classes, sym);
  return check_launch("dg_embed_labels_bwd");
}

extern "C" int dg_gp_interp(const void* labels, int label_bytes, const float* fake, const float* eps, float* out, long long rows,
                            long long rows_per_mol, int classes, void* stream) {
  DG_TRACE("dg_gp_interp", labels, label_bytes, fake, eps, out, rows, rows_per_mol, classes);
  if (rows < 0 || rows_per_mol <= 0 || classes <= 0 || rows % rows_per_mol) return fail("dg_gp_interp: bad shape rows=%lld rows_per_mol=%lld classes=%d", rows, rows_per_mol, classes);
  if (label_bytes != 1 && label_bytes != 8) return fail("dg_gp_interp: labels are uint8 or int64 (label_bytes=%d)", label_bytes);
  if (rows == 0) return 0;
  const long long total = rows * classes;
  const int grid = grid_for(total, 256 * 4);
  if (label_bytes == 1)
    gp_interp_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)labels, fake, eps, out, total, rows_per_mol, classes);
  else
    gp_interp_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, fake, eps, out, total, rows_per_mol, classes);
  return check_launch("dg_gp_interp");
}

extern "C" int dg_gp_penalty(const float* g_node, const float* g_edge, float* penalty, float* coef, float* sq_scratch, int batch,
                             long long len_node, long long len_edge, void* stream) {
  DG_TRACE("dg_gp_penalty", g_node, g_edge, penalty, coef, sq_scratch, batch, len_node, len_edge);
  if (batch <= 0 || len_node < 0 || len_edge < 0) return fail("dg_gp_penalty: bad shape batch=%d", batch);
  gp_sqnorm_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(g_node, g_edge, sq_scratch, len_node, len_edge);
  gp_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sq_scratch, penalty, coef, batch);
  return check_launch("dg_gp_penalty");
}

extern "C" int dg_gp_penalty_bwd(const float* g, const float* coef, const float* upstream, float* out, int batch, long long per_mol,
                                 void* stream) {
  DG_TRACE("dg_gp_penalty_bwd", g, coef, upstream, out, batch, per_mol);
  if (batch <= 0 || per_mol <= 0) return fail("dg_gp_penalty_bwd: bad shape batch=%d per_mol=%lld", batch, per_mol);
  const long long total = (long long)batch * per_mol;
  gp_scale_kernel<<<grid_for(total, 256 * 4), 256, 0, (cudaStream_t)stream>>>(g, coef, upstream, out, total, per_mol);
  return check_launch("dg_gp_penalty_bwd");
}

extern "C" int dg_readout_argmax(const float* x, const float* w, const float* bias, float* logits, void* idx, int idx_bytes,
                                 long long rows, int D, int classes, void* stream) {
  DG_TRACE("dg_readout_argmax", x, w, bias, logits, idx, idx_bytes, rows, D, classes);
  if (rows < 0 || classes <= 0 || classes > kEmbMaxC) return fail("dg_readout_argmax: bad shape rows=%lld classes=%d (classes <= %d)", rows, classes, kEmbMaxC);
  if (D != kEmbD) return fail("dg_readout_argmax: needs D == 128, got %d", D);
  if (idx != nullptr && idx_bytes != 1 && idx_bytes != 8) return fail("dg_readout_argmax: indices are uint8 or int64 (idx_bytes=%d)", idx_bytes);
  if (rows == 0) return 0;
  const int grid = grid_for(rows, 8 * 4 * 4);
  if (idx_bytes == 1)
    readout_argmax_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, logits, (unsigned char*)idx, rows, classes);
  else
    readout_argmax_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, logits, (long long*)idx, rows, classes);
  return check_launch("dg_readout_argmax");
}

extern "C" int dg_adamw_flat(float* p, const float* g, float* m, float* v, const void* segs, int nseg, float lr, float beta1,
                             float beta2, float eps, float weight_decay, void* stream) {
  DG_TRACE("dg_adamw_flat", p, g, m, v, segs, nseg, lr, beta1, beta2, eps, weight_decay);
  if (nseg <= 0 || nseg > 65535) return fail("dg_adamw_flat: segments must be in [1, 65535], got %d", nseg);
  dim3 grid(8, nseg);
  adamw_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (const AdamSeg*)segs, nseg, lr, beta1, beta2, eps, weight_decay);
  return check_launch("dg_adamw_flat");
}

extern "C" int dg_label2onehot(const void* labels, int label_bytes, float* out, long long n, int classes, void* stream) {
  DG_TRACE("dg_label2onehot", labels, label_bytes, out, n, classes);
  if (n < 0 || classes <= 0) return fail("dg_label2onehot: bad shape n=%lld classes=%d", n, classes);
  if (label_bytes != 1 && label_bytes != 8) return fail("dg_label2onehot: labels are uint8 or int64 (label_bytes=%d)", label_bytes);
  if (n == 0) return 0;
  const long long total = n * classes;
  const int grid = grid_for(total, 256);
  if (label_bytes == 1)
    onehot_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)labels, out, total, classes);
  else
    onehot_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, out, total, classes);
  return check_launch("dg_label2onehot");
}

extern "C" int dg_argmax_last(const float* x, long long* out, long long rows, int C, void* stream) {
  DG_TRACE("dg_argmax_last", x, out, rows, C);
  if (rows < 0 || C <= 0) return fail("dg_argmax_last: bad shape rows=%lld C=%d", rows, C);
  if (rows == 0) return 0;
  argmax_last_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(x, out, rows, C);
  return check_launch("dg_argmax_last");
}

// ---- the producer of the path's inputs: load_molecules, src/data/utils.py:128-143 ---------------
// torch_geometric.utils.to_dense_adj (PyG 2.2.0, the reference's pinned dependency; not vendored) restated:
//   num_nodes[b] = #{v : batch[v] = b};  cum = exclusive prefix sum;  for every edge (s, t) with attribute a (1 when absent):
//   adj[batch[s], s - cum[batch[s]], t - cum[batch[t]]] += a, edges whose local index reaches max_num_nodes dropped.
// Integer scatter-add into [B,N,N] int32 (duplicate edges add, as PyG's scatter(reduce='add')), then narrowed to the 1-byte label
// wire format with the class-range check label2onehot's scatter_ would make.  HBM-bound integer work: one pass over the edge list.
namespace dg {

__global__ void count_nodes_kernel(const long long* __restrict__ batch, unsigned long long* __restrict__ counts, long long V, int B) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const long long b = batch[v];
    if (b < 0 || b >= B) atomicOr(&g_label_error, 2);
    else atomicAdd(&counts[b + 1], 1ull);
  }
}
// in-place inclusive scan of counts[1..B] (counts[0] = 0) -> cum[b] = first node of graph b; one block, B is a batch size
__global__ void __launch_bounds__(1024) scan_nodes_kernel(unsigned long long* __restrict__ cum, int B) {
  __shared__ unsigned long long part[1024];
  const int per = (B + 1023) / 1024, lo = 1 + threadIdx.x * per, hi = min(B + 1, lo + per);
  unsigned long long s = 0;
  for (int i = lo; i < hi; ++i) s += cum[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; ++i) { const unsigned long long t = part[i]; part[i] = run; run += t; }
  }
  __syncthreads();
  unsigned long long run = part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { run += cum[i]; cum[i] = run; }
}
__global__ void dense_adj_kernel(const long long* __restrict__ edge_index, const long long* __restrict__ batch,
                                 const long long* __restrict__ edge_attr, const unsigned long long* __restrict__ cum,
                                 int* __restrict__ adj, long long E, long long V, int B, int N) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long s = edge_index[e], t = edge_index[E + e];
    if (s < 0 || s >= V || t < 0 || t >= V) { atomicOr(&g_label_error, 2); continue; }
    const long long bs = batch[s], bt = batch[t];
    if (bs < 0 || bs >= B || bt < 0 || bt >= B) continue;          // (flagged by count_nodes_kernel)
    const long long i1 = s - (long long)cum[bs], i2 = t - (long long)cum[bt];
    if (i1 >= N || i2 >= N) continue;                                // to_dense_adj's max_num_nodes mask
    atomicAdd(&adj[(bs * N + i1) * N + i2], edge_attr ? (int)edge_attr[e] : 1);
  }
}
__global__ void narrow_labels_kernel(const int* __restrict__ adj, unsigned char* __restrict__ out, long long n, int classes) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int a = adj[i];
    if (a < 0 || a >= classes) atomicOr(&g_label_error, 1);          // label2onehot's scatter_ would raise on it
    out[i] = (unsigned char)a;
  }
}

// ---- SNN / internal-diversity metric: average_agg_tanimoto, src/util/utils.py:566-611 ------------------------
// The reference multiplies 0/1 fingerprint matrices as fp32 (torch.mm) to count common bits.  Here fingerprints are bit-packed
// (F/64 words of 64 bits), tp = popcount(x & y) -- exact integers, as the fp32 GEMM's are below 2^24 -- and
// jac = tp / (|x| + |y| - tp) is ONE IEEE fp32 division of the same integers (0/0 -> 1 as the reference's NaN patch): bit-exact.
// Thread = one generated fingerprint in registers; the stock fingerprints pass through shared memory as broadcast reads.
constexpr int kTanWords = 32;          // up to 2048-bit fingerprints
constexpr int kTanTile = 128;          // stock fingerprints per shared-memory tile

template <typename T>
__global__ void pack_bits_kernel(const T* __restrict__ vecs, unsigned long long* __restrict__ bits, int* __restrict__ cnt,
                                 long long rows, int F, int words) {
  const long long total = rows * words;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / words;
    const int w = (int)(i - r * words);
    unsigned long long word = 0;
    const int nb = min(64, F - w * 64);
    for (int k = 0; k < nb; ++k) word |= (unsigned long long)(vecs[r * F + w * 64 + k] != (T)0) << k;
    bits[i] = word;
    if (word) atomicAdd(&cnt[r], __popcll(word));
  }
}

template <int W>
__global__ void __launch_bounds__(128) tanimoto_agg_kernel(const unsigned long long* __restrict__ stock, const int* __restrict__ stock_cnt,
                                                           long long S, const unsigned long long* __restrict__ gen,
                                                           const int* __restrict__ gen_cnt, long long G, int agg, float p,
                                                           float* __restrict__ out_max, double* __restrict__ out_sum, int s_chunk) {
  __shared__ unsigned long long sS[kTanTile * W];
  __shared__ int sC[kTanTile];
  const long long g = (long long)blockIdx.x * 128 + threadIdx.x;
  unsigned long long y[W];
#pragma unroll
  for (int w = 0; w < W; ++w) y[w] = g < G ? gen[g * W + w] : 0ull;
  const int yc = g < G ? gen_cnt[g] : 0;
  const long long s_lo = (long long)blockIdx.y * s_chunk, s_hi = min(S, s_lo + s_chunk);
  float best = 0.f;                                          // (agg_tanimoto starts from zeros: utils.py:584)
  double sum = 0.0;
  for (long long s0 = s_lo; s0 < s_hi; s0 += kTanTile) {
    const int ns = (int)min((long long)kTanTile, s_hi - s0);
    __syncthreads();
    for (int i = threadIdx.x; i < ns * W; i += 128) sS[i] = stock[s0 * W + i];
    for (int i = threadIdx.x; i < ns; i += 128) sC[i] = stock_cnt[s0 + i];
    __syncthreads();
    float part = 0.f;
    for (int s = 0; s < ns; ++s) {
      int tp = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) tp += __popcll(sS[s * W + w] & y[w]);
      const int den = sC[s] + yc - tp;
      float jac = den == 0 ? 1.f : __fdiv_rn((float)tp, (float)den);      // utils.py:596-597 (NaN -> 1)
      if (agg == 0) best = fmaxf(best, jac);
      else part += (p == 1.f ? jac : powf(jac, p));                          // utils.py:598-599, :604
    }
    sum += (double)part;
  }
  if (g < G) {
    if (agg == 0) atomicMax(reinterpret_cast<int*>(out_max) + g, __float_as_int(best));   // (non-negative floats order as ints)
    else atomicAdd(out_sum + g, sum);
  }
}

}  // namespace dg

extern "C" int dg_to_dense_adj(const long long* edge_index, const long long* batch, const long long* edge_attr, int* adj,
                               unsigned long long* cum_nodes, long long E, long long V, int B, int N, void* stream) {
  if (E < 0 || V < 0 || B <= 0 || N <= 0) return fail("dg_to_dense_adj: bad shape E=%lld V=%lld B=%d N=%d", E, V, B, N);
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(adj, 0, (size_t)B * N * N * sizeof(int), s) != cudaSuccess ||
      cudaMemsetAsync(cum_nodes, 0, (size_t)(B + 1) * sizeof(unsigned long long), s) != cudaSuccess)
    return fail("dg_to_dense_adj: memset failed");
  if (V > 0) count_nodes_kernel<<<grid_for(V, 256), 256, 0, s>>>(batch, cum_nodes, V, B);
  scan_nodes_kernel<<<1, 1024, 0, s>>>(cum_nodes, B);
  if (E > 0) dense_adj_kernel<<<grid_for(E, 256), 256, 0, s>>>(edge_index, batch, edge_attr, cum_nodes, adj, E, V, B, N);
  return check_launch("dg_to_dense_adj");
}

extern "C" int dg_narrow_labels(const int* adj, unsigned char* out, long long n, int classes, void* stream) {
  if (n < 0 || classes <= 0 || classes > 256) return fail("dg_narrow_labels: bad shape n=%lld classes=%d", n, classes);
  if (n == 0) return 0;
  narrow_labels_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(adj, out, n, classes);
  return check_launch("dg_narrow_labels");
}

extern "C" int dg_pack_bits(const void* vecs, int elem_bytes, unsigned long long* bits, int* popcounts, long long rows, int F,
                            void* stream) {
  if (rows < 0 || F <= 0) return fail("dg_pack_bits: bad shape rows=%lld F=%d", rows, F);
  if (elem_bytes != 1 && elem_bytes != 4) return fail("dg_pack_bits: elements are uint8 or float32 (elem_bytes=%d)", elem_bytes);
  if (rows == 0) return 0;
  const int words = (F + 63) / 64;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(popcounts, 0, (size_t)rows * sizeof(int), s) != cudaSuccess) return fail("dg_pack_bits: memset failed");
  if (elem_bytes == 1)
    pack_bits_kernel<unsigned char><<<grid_for(rows * words, 256), 256, 0, s>>>((const unsigned char*)vecs, bits, popcounts, rows, F, words);
  else
    pack_bits_kernel<float><<<grid_for(rows * words, 256), 256, 0, s>>>((const float*)vecs, bits, popcounts, rows, F, words);
  return check_launch("dg_pack_bits");
}

extern "C" int dg_tanimoto_agg(const unsigned long long* stock_bits, const int* stock_cnt, long long S,
                               const unsigned long long* gen_bits, const int* gen_cnt, long long G, int words, int agg, float p,
                               float* out_max, double* out_sum, void* stream) {
  if (S < 0 || G < 0 || words <= 0 || words > kTanWords) return fail("dg_tanimoto_agg: bad shape S=%lld G=%lld words=%d (<= %d)", S, G, words, kTanWords);
  if (agg != 0 && agg != 1) return fail("dg_tanimoto_agg: agg is 0 (max) or 1 (sum)");
  if ((agg == 0 ? (void*)out_max : (void*)out_sum) == nullptr) return fail("dg_tanimoto_agg: the output of this aggregation is NULL");
  if (S == 0 || G == 0) return 0;
  // enough CTAs for the machine: split the stock set when there are few generated fingerprints
  const long long gx = (G + 127) / 128;
  long long sy = (2LL * sm_count() + gx - 1) / gx;
  const long long max_sy = (S + kTanTile - 1) / kTanTile;
  if (sy > max_sy) sy = max_sy;
  if (sy < 1) sy = 1;
  if (sy > 65535) sy = 65535;
  long long chunk = (S + sy - 1) / sy;
  chunk = (chunk + kTanTile - 1) / kTanTile * kTanTile;
  sy = (S + chunk - 1) / chunk;
  dim3 grid((unsigned)gx, (unsigned)sy);
  cudaStream_t s = (cudaStream_t)stream;
#define DG_TAN(Wn) tanimoto_agg_kernel<Wn><<<grid, 128, 0, s>>>(stock_bits, stock_cnt, S, gen_bits, gen_cnt, G, agg, p, out_max, out_sum, (int)chunk)
  switch (words) {
    case 1: DG_TAN(1); break;
    case 2: DG_TAN(2); break;
    case 4: DG_TAN(4); break;
    case 8: DG_TAN(8); break;
    case 16: DG_TAN(16); break;
    case 32: DG_TAN(32); break;
    default: return fail("dg_tanimoto_agg: words must be a power of two <= %d (pad the fingerprints with zero words), got %d", kTanWords, words);
  }
#undef DG_TAN
  return check_launch("dg_tanimoto_agg");
}

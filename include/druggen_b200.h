/* druggen_b200 -- C-ABI of the B200 (sm_100a) graph-transformer encoder hot path.
 *
 * The reference (HUBioDataLab/DrugGEN) has no FFI of its own: its hot path is reached through the
 * Python class API of src/model/layers.py.  Each entry point below replaces the ATen kernels that
 * one group of reference lines dispatches; the file:line it replaces is cited per function.
 * INTEGRATION.md shows the ctypes binding (druggen_b200/_lib.py) a maintainer drops in.
 *
 * Conventions
 *   - all tensors are fp32, row-major, contiguous, device pointers on the current device;
 *   - the caller owns every buffer (outputs and accumulators included); kernels never allocate;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous, no hidden syncs;
 *   - return value 0 = launched; non-zero = rejected (dg_last_error() has the reason), nothing ran;
 *   - B molecules, N atoms per molecule ("vertexes"), D channels ("dim"); "rows" are the flattened
 *     node rows [B*N] or edge rows [B*N*N];
 *   - `prec` selects the arithmetic of the dense contractions only:
 *       DG_PREC_FP32   CUDA-core fp32 FMA                     (parity mode)
 *       DG_PREC_BF16   tcgen05.mma kind::f16 bf16 x bf16 -> fp32 in TMEM   (throughput mode)
 *       DG_PREC_BF16X3 tcgen05 split precision: every fp32 operand x = hi + lo (two bf16), a.b = a_hi b_hi + a_hi b_lo + a_lo b_hi
 *                      accumulated in fp32 in TMEM -- 16 operand mantissa bits on the tensor cores (tensor-core parity mode);
 *                      all tensors stay fp32 in HBM, shapes wider than 128 run as 128-wide slices
 *     every other operation (LayerNorm, softmax, modulation, reductions) is fp32 in all modes.
 */
#ifndef DRUGGEN_B200_H
#define DRUGGEN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DG_ABI_VERSION 6
#define DG_PREC_FP32 0
#define DG_PREC_BF16 1
#define DG_PREC_BF16X3 2

int dg_abi_version(void);
/* Thread-local text of the last rejected call. */
const char* dg_last_error(void);
/* 1 when the tcgen05 paths were compiled in and the current device is sm_100. */
int dg_has_tcgen05(void);

/* Run-time options (process-wide; read at launch time).  Returns 0 / the value. */
#define DG_OPT_L2_PREFETCH 0 /* bit mask: which kernels issue bulk L2 prefetches (cp.async.bulk.prefetch.L2) ahead of their
                              * loads.  Default DG_PF_GEMM_TN | DG_PF_ATTN_FWD: the kernels where it measured faster on B200. */
#define DG_PF_CHAIN 1      /* dg_mlp_*, dg_attn_edge_fwd */
#define DG_PF_ROWS_GEMM 2  /* dg_rows_gemm */
#define DG_PF_GEMM_TN 4    /* dg_gemm_tn with M = N = 128 */
#define DG_PF_ATTN_FWD 8   /* dg_attn_scores_fwd, dg_softmax_agg16_fwd */
#define DG_PF_ATTN_BWD 16  /* dg_attn_scores_bwd */
#define DG_PF_GEMM_TN_WIDE 32 /* dg_gemm_tn with M or N > 128 */
#define DG_PF_CHAIN_KEEP 64 /* chain kernels: the loader's reads of x carry L2::evict_last (x is re-read as the residual) */
#define DG_OPT_ATTN_BWD 1    /* dg_attn_scores_bwd with the forward's statistics: 0 (default) = the TMA-fed ring kernel where it applies
                              * (N <= 48; 8 warps x <= 6 key atoms), 2 = the ring kernel with 16 warps x <= 3 key atoms,
                              * 1 = always the 4-warp register-staged kernel (A/B switches) */
#define DG_PF_SECOND 128   /* second-order attention kernels (dg_modulate_bwd*, dg_softmax_agg_bwd*): next query atom's rows.  Off by
                              * default: measured 20-27 % slower on the modulate kernels, neutral on the softmax ones (B200) */
#define DG_OPT_COUNT 2
int dg_set_option(int key, int value);
int dg_get_option(int key);

/* ---- dense contractions ------------------------------------------------------------------ */
/* out[R,N] = epi(a[R,K] . op(w) + bias) + resid;  w is [N,K] when w_is_nk (nn.Linear layout) else [K,N].
 * epi: optional ReLU, optional gate: out *= (gate[R,N] > 0) (ReLU backward); resid[R,N] optional
 * (gradient accumulation / residual fused into the store).
 * Replaces nn.Linear forward / addmm (layers.py:51,53,111-113,116,127,135) and the dgrad `mm`s. */
int dg_rows_gemm(const void* a, const float* w, int w_is_nk, const float* bias, int relu,
                 const void* gate, const float* resid, void* out, long long R, int K, int N, int prec,
                 int flags, void* stream);
/* storage flags (tensor-core mode only): tensors that are consumed solely as contraction operands / sign
 * masks may live in HBM as bf16 -- the contraction rounds them to bf16 anyway, so no accuracy is lost */
#define DG_A_BF16 1     /* rows_gemm: a is bf16      | gemm_tn: a is bf16 */
#define DG_OUT_BF16 2   /* rows_gemm: out is bf16    | gemm_tn: b is bf16 */
#define DG_GATE_BF16 4  /* rows_gemm: gate is bf16 */
/* out[M,N] += a[R,M]^T . b[R,N]   (split over rows, accumulated atomically: zero `out` first
 * unless accumulating); colsum_a[M] += column sums of a (the bias gradient, same pass), optional.
 * Replaces the weight-gradient `mm`s and bias-gradient `sum`s of autograd. */
int dg_gemm_tn(const void* a, const void* b, float* out, float* colsum_a, long long R, int M, int N,
               int prec, int flags, void* stream);
/* out[N] += column sums of a[R,N]   (bias gradients). */
int dg_colsum(const float* a, float* out, long long R, int N, void* stream);
/* out = x * (ref > 0), n elements   (threshold_backward of layers.py:52). */
int dg_gate_mul(const float* x, const float* ref, float* out, long long n, void* stream);

/* ---- residual + LayerNorm (eps inside the sqrt, affine) : layers.py:185,187-192 -------------- */
/* out = LN(a + b) * gamma + beta over D; b may be NULL. */
int dg_add_ln_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* out,
                  long long R, int D, float eps, void* stream);
/* dz[R,D] written (accumulate != 0: dz += , the add of another path's cotangent rides in the store);
 * dgamma[D] += , dbeta[D] +=   (z = a + b recomputed). */
int dg_add_ln_bwd(const float* dy, const float* a, const float* b, const float* gamma, float* dz,
                  float* dgamma, float* dbeta, long long R, int D, float eps, int accumulate, void* stream);
/* Second order: gradient of <u,dz> + <vg,dgamma> + <vb,dbeta> w.r.t. dy, z, gamma.
 * vg, vb may be NULL.  g_gamma[D] += . */
int dg_add_ln_bwd_bwd(const float* u, const float* vg, const float* vb, const float* dy, const float* a,
                      const float* b, const float* gamma, float* g_dy, float* g_z, float* g_gamma,
                      long long R, int D, float eps, void* stream);

/* ---- edge-modulated scores : layers.py:119-125 ---------------------------------------------- */
/* A[b,i,j,:] = c * q[b,i,:] * k[b,j,:] * (e^2 + e)[b,i,j,:] */
int dg_modulate_fwd(const float* q, const float* k, const float* e, float c, float* out, int B, int N, int D,
                    void* stream);
/* de written; dq, dk += (zero first). */
int dg_modulate_bwd(const float* da, const float* q, const float* k, const float* e, float c, float* dq,
                    float* dk, float* de, int B, int N, int D, void* stream);
/* g_da, g_e written; g_q, g_k += (zero first). */
int dg_modulate_bwd_bwd(const float* uq, const float* uk, const float* ue, const float* da, const float* q,
                        const float* k, const float* e, float c, float* g_da, float* g_q, float* g_k,
                        float* g_e, int B, int N, int D, void* stream);

/* ---- softmax over key atoms + value aggregation : layers.py:130-134 ------------------------- */
/* g[b,i,:] = sum_j softmax_j(a[b,i,j,:]) * v[b,j,:] */
int dg_softmax_agg_fwd(const float* a, const float* v, float* g, int B, int N, int D, void* stream);
/* da written (da += when `accumulate`); dv += (zero first). */
int dg_softmax_agg_bwd(const float* dg, const float* a, const float* v, float* da, float* dv, int accumulate,
                       int B, int N, int D, void* stream);
/* g_dg, g_a written; g_v += (zero first). */
int dg_softmax_agg_bwd_bwd(const float* ua, const float* uv, const float* dg, const float* a, const float* v,
                           float* g_dg, float* g_a, float* g_v, int B, int N, int D, void* stream);

/* ---- fused attention scores (fp32; D == 128, N >= 4) : layers.py:119-125 + :130-134 in one pass ------------ */
/* a[b,i,j,:] = c q_i k_j (e^2+e) written (operand of out_e) and g[b,i,:] = sum_j softmax_j(a) v_j; e read once. */
int dg_attn_scores_fwd(const float* q, const float* k, const float* v, const float* e, float c, float* a,
                       float* g, float* stat_m, float* stat_inv, int B, int N, int D, void* stream);
/* (stat_m, stat_inv: optional [B,N,D] outputs = per-channel softmax max and 1/sum, for the backward.)
 * First-order backward of the pair: de written from dg (softmax path) + da_in (out_e path, may be NULL);
 * dq, dk, dv: zero all three first and do not count on accumulation (dq is reduced into; dk / dv are reduced into by the
 * 4-warp kernel and STORED by the ring kernel, whose CTA owns all rows of a molecule).  Scores are recomputed from e, q, k; with (stat_m, stat_inv, g) from the
 * forward the statistics sweep is skipped (and the warps of a molecule never synchronise), with NULLs it is redone. */
int dg_attn_scores_bwd(const float* dg, const float* da_in, const float* q, const float* k, const float* v,
                       const float* e, float c, const float* stat_m, const float* stat_inv, const float* g,
                       void* de, float* dq, float* dk, float* dv, int B, int N, int D, int de_bf16, void* stream);
/* (de_bf16 bit 0: de is written as bf16 [B,N,N,D] -- it is only ever a contraction operand (dWe, dy), so in the
 * tensor-core mode nothing is lost.  bit 1: the recomputed scores are rounded to bf16 before the softmax term, matching
 * statistics that were taken from the bf16 scores of dg_attn_edge_fwd / dg_softmax_agg16_fwd.  bit 2: da_in is bf16
 * [B,N,N,D] -- the out_e path's gradient as dg_rows_gemm stores it with a bf16 output.  bit 3: de (fp32) += instead of
 * being written.)
 * g and the statistics from bf16 scores a[B,N,N,D] (the side output of dg_attn_edge_fwd) -- the forward's
 * softmax-aggregate (layers.py:130-134) at 256 B per edge row. */
int dg_softmax_agg16_fwd(const void* a_bf16, const float* v, float* g, float* stat_m, float* stat_inv, int B, int N,
                         int D, void* stream);

/* ---- fused tcgen05 kernels (bf16 operands, fp32 accumulate / epilogue) ----------------------------- */
/* out = LN(x + fc2(relu(fc1(x) + b1)) + b2) * gamma + beta   -- the whole residual MLP of one stream in one
 * kernel (layers.py:41-54 + :191 / :192); x,out:[R,D] with D == 128, hidden width H in {128,256,384}.
 * workspace: >= 2 * (H/128) * 32768 bytes, 128-byte aligned (bf16 swizzled weight stages are packed there). */
int dg_mlp_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
               const float* gamma, const float* beta, float* out, long long R, int D, int H, float eps,
               void* workspace, long long workspace_bytes, void* stream);

/* First half of the residual-MLP backward in one kernel: recompute h = relu(fc1(x)+b1) and z = x + fc2(h) + b2,
 * then the LayerNorm backward of `dout` through z.  Writes dz[R,D] (fp32); optional side outputs (NULL to skip):
 * h_bf16[R,H] (bf16: the operand of the fc2 weight gradient) and relu_mask[R][H/64] (uint64, 8-byte aligned: bit i of
 * word w = (h[w*64+i] > 0) -- all that dg_mlp_bwd_dgrad needs of h, H/8 bytes per row instead of 2H);
 * dgamma[D], dbeta[D] += (zero first; pass NULL for both to skip the column sums -- dgrad-only passes).  Same workspace
 * contract as dg_mlp_fwd. */
int dg_mlp_bwd_ln(const float* x, const float* dout, const float* w1, const float* b1, const float* w2,
                  const float* b2, const float* gamma, float* dz, void* h_bf16, void* relu_mask, float* dgamma,
                  float* dbeta, long long R, int D, int H, float eps, void* workspace, long long workspace_bytes,
                  void* stream);
/* Second half: dh = (dz . W2) * (h > 0) written as bf16 [R,H] (dh_bf16, NULL to skip: only the fc1 weight gradient reads
 * it); dx = dz + dh . W1 written [R,D] (fp32).  The sign of h comes from relu_mask (dg_mlp_bwd_ln's layout) when given,
 * else from h_bf16.  (threshold_backward + the two dgrad `mm`s + the residual add of layers.py:51-54,191-192.) */
int dg_mlp_bwd_dgrad(const float* dz, const void* h_bf16, const void* relu_mask, const float* w1, const float* w2,
                     float* dx, void* dh_bf16, long long R, int D, int H, void* workspace, long long workspace_bytes,
                     void* stream);

/* The edge half of the attention block in one tcgen05 kernel (layers.py:116,123-127 + the residual and LayerNorm of
 * :188/:190):  E = y.We^T + be;  A = c q_i k_j (E^2 + E);  out = LN4(y + A.Woe^T + boe) * gamma + beta.
 * y,out:[B*N*N,D] fp32 (D == 128), q,k:[B,N,D] fp32.  Optional side outputs (NULL to skip): a_bf16 [B*N*N,D] bf16
 * (scores: operand of dg_softmax_agg16_fwd and of the out_e weight gradient), e_out [B*N*N,D] fp32 and z_out
 * [B*N*N,D] fp32 = y + out_e(A), the input of LN4 (both for the backward).  E and A never reach HBM as fp32
 * unless asked for.  workspace: >= 65536 bytes, 128-byte aligned. */
int dg_attn_edge_fwd(const float* y, const float* q, const float* k, const float* we, const float* be,
                     const float* woe, const float* boe, const float* gamma, const float* beta, float c,
                     float* out, void* a_bf16, float* e_out, float* z_out, int B, int N, int D, float eps,
                     void* workspace, long long workspace_bytes, void* stream);

/* ---- either side of the encoder path -------------------------------------------------------- */
/* src/data/utils.py:15-23 label2onehot: out[n, classes] fp32 = one-hot of labels[n]; labels are int64 (label_bytes 8, what
 * the reference's to_dense_adj produces) or uint8 (label_bytes 1: a 1-byte-per-edge wire format for the host->device copy).
 * A label outside [0, classes) raises the device flag that dg_label_error() reads (torch's scatter_ raises); its row is all zero. */
int dg_label2onehot(const void* labels, int label_bytes, float* out, long long n, int classes, void* stream);
/* Reads (and with `clear` resets) the out-of-range-label flag set by dg_label2onehot / dg_embed_labels_* / dg_gp_interp since the
 * last clear.  Synchronises with the device.  1 = some label was outside [0, classes); 0 = none; -1 = the flag could not be read. */
int dg_label_error(int clear);
/* inference.py:197-198 torch.max(t, -1)[1]: out[rows] int64 = index of the first maximum of each row of x[rows, C]
 * (a NaN wins; the first NaN), i.e. ATen's CPU result, bit for bit. */
int dg_argmax_last(const float* x, long long* out, long long rows, int C, void* stream);
/* models.py:94 / :199 for dense inputs: out[b,i,j,:] = (e[b,i,j,:] + e[b,j,i,:]) / 2 in one pass ([B,N,N,D] fp32, D % 4 == 0,
 * out != e).  Self-adjoint: the backward is the same call on the gradient. */
int dg_symmetrize(const float* e, float* out, int B, int N, int D, void* stream);

/* models.py:91-94 / 196-199 for one-hot inputs given as labels: y[r,:] = lut[labels[r],:] (node rows, sym = 0) or, for edge rows
 * r = (b n + i) n + j with sym = 1, y[r,:] = (lut[a_ij,:] + lut[a_ji,:]) / 2 -- the prologue MLP of a one-hot row is a row of the
 * [classes, D] table lut = act(W2 act(W1[:,l] + b1) + b2) the host computes from the weights; D == 128, classes <= 16. */
int dg_embed_labels_fwd(const void* labels, int label_bytes, const float* lut, float* y, long long rows, int n, int classes,
                        int D, int sym, void* stream);
/* its backward: dlut[l,:] += sum_r dy[r,:] ([a_ij = l] + [a_ji = l]) / 2   (sym = 0: sum over rows with labels[r] = l); zero first. */
int dg_embed_labels_bwd(const void* labels, int label_bytes, const float* dy, float* dlut, long long rows, int n, int classes,
                        int D, int sym, void* stream);
/* loss.py:21-26 with the real side given as labels: out[r,c] = eps[b] * [labels[r] == c] + (1 - eps[b]) * fake[r,c],
 * b = r / rows_per_mol; rounded exactly as torch's mul / rsub / mul / add kernels round (no FMA contraction). */
int dg_gp_interp(const void* labels, int label_bytes, const float* fake, const float* eps, float* out, long long rows,
                 long long rows_per_mol, int classes, void* stream);
/* loss.py:42-47: penalty[0] = mean_b (|g_b| - 1)^2 with |g_b| the L2 norm over concat(g_node[b,:], g_edge[b,:]);
 * coef[b] = 2 (|g_b| - 1) / (batch |g_b|) for the backward; sq_scratch[batch] is workspace. */
int dg_gp_penalty(const float* g_node, const float* g_edge, float* penalty, float* coef, float* sq_scratch, int batch,
                  long long len_node, long long len_edge, void* stream);
/* d penalty / d g: out[b,i] = upstream[0] * coef[b] * g[b,i]   (g = g_node or g_edge, per_mol elements per molecule). */
int dg_gp_penalty_bwd(const float* g, const float* coef, const float* upstream, float* out, int batch, long long per_mol,
                      void* stream);
/* models.py:100-101 + inference.py:197-198 in one pass over x[rows, D]: logits[rows, classes] = x w^T + bias (NULL to skip) and
 * idx[rows] = first maximal class as int64 (idx_bytes 8) or uint8 (idx_bytes 1) (NULL to skip); D == 128, classes <= 16. */
int dg_readout_argmax(const float* x, const float* w, const float* bias, float* logits, void* idx, int idx_bytes, long long rows,
                      int D, int classes, void* stream);
/* train.py:213-214 torch.optim.AdamW over ONE flat buffer per network: p, g, m (exp_avg), v (exp_avg_sq) are parallel fp32 arrays;
 * segs[nseg] (device) = { int64 begin, end; float bias_correction1, sqrt(bias_correction2); int32 active, pad } per parameter tensor:
 * inactive tensors (gradient None) are untouched, exactly as torch skips them. */
int dg_adamw_flat(float* p, const float* g, float* m, float* v, const void* segs, int nseg, float lr, float beta1, float beta2,
                  float eps, float weight_decay, void* stream);

/* ---- the producer of the path's inputs and the evaluation metric ------------------------ */
/* load_molecules, src/data/utils.py:128-143: torch_geometric.utils.to_dense_adj (PyG 2.2.0, the reference's pinned dependency)
 * on the device.  edge_index [2,E] int64 (row 0 sources, row 1 targets, global node ids), batch [V] int64 (graph of each node,
 * sorted), edge_attr [E] int64 bond labels or NULL (ones) -> adj [B,N,N] int32, zeroed here: adj[batch[s], s - first(batch[s]),
 * t - first(batch[t])] += attr, edges whose local index is >= N dropped (max_num_nodes), duplicate edges add.
 * cum_nodes: workspace of (B + 1) uint64 (receives the first node id of every graph).  A node or graph id out of range raises the
 * dg_label_error() flag (bit 2). */
int dg_to_dense_adj(const long long* edge_index, const long long* batch, const long long* edge_attr, int* adj,
                    unsigned long long* cum_nodes, long long E, long long V, int B, int N, void* stream);
/* int32 labels -> the 1-byte label wire format (input of dg_embed_labels_fwd / dg_label2onehot / dg_gp_interp); a value outside
 * [0, classes) raises the dg_label_error() flag, as label2onehot's scatter_ (src/data/utils.py:21) would raise on it. */
int dg_narrow_labels(const int* adj, unsigned char* out, long long n, int classes, void* stream);
/* average_agg_tanimoto, src/util/utils.py:566-611, on bit-packed fingerprints.  dg_pack_bits: vecs [rows,F] of uint8 (elem_bytes 1)
 * or float32 (elem_bytes 4), non-zero = bit set -> bits [rows, ceil(F/64)] uint64 (bit k of word w = element 64 w + k) and
 * popcounts [rows] int32. */
int dg_pack_bits(const void* vecs, int elem_bytes, unsigned long long* bits, int* popcounts, long long rows, int F, void* stream);
/* For every generated fingerprint g: jac(s,g) = |s & g| / (|s| + |g| - |s & g|) over all S stock fingerprints (0/0 -> 1, utils.py:597;
 * the division is one IEEE fp32 division of exact integers: bit-equal to the reference's fp32 GEMM + division), aggregated as
 * agg 0: out_max[g] = max(out_max[g], max_s jac)      (float32 [G]; start from zeros, utils.py:584)
 * agg 1: out_sum[g] += sum_s jac^p                    (float64 [G]; the caller divides by the count, utils.py:604-607)
 * words = uint64 words per fingerprint: a power of two <= 32 (pad with zero words). */
int dg_tanimoto_agg(const unsigned long long* stock_bits, const int* stock_cnt, long long S, const unsigned long long* gen_bits,
                    const int* gen_cnt, long long G, int words, int agg, float p, float* out_max, double* out_sum, void* stream);

/* ---- block-level entry points (dg_block_fwd / dg_block_bwd) ------------------------------------------------------
 * One host call per direction of an encoder block, Encoder_Block.forward (layers.py:174-193) and its autograd backward, in the
 * tensor-core throughput mode (DG_PREC_BF16; D == 128, H in {128,256,384}, 4 <= N <= 212): a fixed sequence of the launches
 * above on the caller's stream over caller-owned buffers -- no allocation, no synchronisation, capturable into a CUDA graph.
 * `params`: the block's DG_BLOCK_PARAMS device pointers in the reference's state-dict order: ln1.{weight,bias},
 * attn.{q,k,v,e,out_e,out_n}.{weight,bias}, ln3.*, ln4.*, mlp.{fc1,fc2}.*, mlp2.{fc1,fc2}.*, ln5.*, ln6.*.
 * `io`: DG_BLK_COUNT device pointers (NULL where a call does not need the buffer).  Node-sized buffers are [B*N, D] fp32,
 * edge-sized ones [B*N*N, D] fp32 unless noted.  workspace: >= 2 * (H/128) * 32768 bytes, 128-byte aligned. */
#define DG_BLOCK_PARAMS 30
enum {
  DG_BLK_X = 0,    /* in : block input x */
  DG_BLK_Y,        /* in : block input y (edge) */
  DG_BLK_X_OUT,    /* fwd out */
  DG_BLK_Y_OUT,    /* fwd out (edge); unused without DG_BLKF_EDGE_OUT */
  DG_BLK_X1,       /* LN1(x)                                   fwd: scratch or kept;  bwd: kept input or scratch (recomputed) */
  DG_BLK_Q, DG_BLK_K, DG_BLK_V,      /* projections of x1     (same) */
  DG_BLK_G,        /* softmax-aggregate g                      (same; with DG_BLKF_STATS the backward reads the forward's) */
  DG_BLK_ON,       /* out_n(g)                                 (same) */
  DG_BLK_X3,       /* LN3(x1 + on)                             (same) */
  DG_BLK_STAT_M, DG_BLK_STAT_INV,    /* softmax max and 1/sum per (molecule, query atom, channel); fwd: written with DG_BLKF_STATS */
  DG_BLK_Y3,       /* LN4(y + out_e(A)) (edge)                 fwd: scratch or kept; without a live edge output: scratch for the scores */
  DG_BLK_A16,      /* the scores A, bf16 (edge, 2 bytes)       (same) */
  DG_BLK_E,        /* E = e(y) (edge)                          fwd: written with DG_BLKF_KEEP, and by a block without edge output */
  DG_BLK_Z4,       /* y + out_e(A) (edge)                      fwd: written with DG_BLKF_KEEP */
  DG_BLK_DXO,      /* bwd in : d x_out (NULL = zeros) */
  DG_BLK_DYO,      /* bwd in : d y_out (NULL = the edge output has no consumer) */
  DG_BLK_DX,       /* bwd out */
  DG_BLK_DY,       /* bwd out (edge) */
  DG_BLK_N_DZ, DG_BLK_N_DX3, DG_BLK_N_DZ3, DG_BLK_N_DG, DG_BLK_N_DQ, DG_BLK_N_DK, DG_BLK_N_DV, DG_BLK_N_T0, DG_BLK_N_T1, /* bwd scratch, node */
  DG_BLK_N_H,      /* bwd scratch [B*N, H] bf16 (only with weight gradients) */
  DG_BLK_N_MASK,   /* bwd scratch [B*N, H/64] uint64 */
  DG_BLK_E_A, DG_BLK_E_B,            /* bwd scratch, edge fp32 (only with a live edge output) */
  DG_BLK_E_H,      /* bwd scratch, bf16: [B*N*N, H] with weight gradients and a live edge output, else [B*N*N, D] */
  DG_BLK_E_MASK,   /* bwd scratch [B*N*N, H/64] uint64 (only with a live edge output) */
  DG_BLK_SCRATCH,  /* bwd scratch, 2 D floats */
  /* ---- dg_block_bwd_bwd only */
  DG_BLK_UX,       /* in : cotangent of dx */
  DG_BLK_UY,       /* in : cotangent of dy (edge) */
  DG_BLK_C_X,      /* out: cotangent of x */
  DG_BLK_C_Y,      /* out: cotangent of y (edge) */
  DG_BLK_C_DXO,    /* out: cotangent of d x_out */
  DG_BLK_C_DYO,    /* out: cotangent of d y_out (edge; only with a live edge output) */
  DG_BLK_N_ARENA,  /* scratch: DG_BLK_BB_NODE_SLOTS node-sized fp32 buffers, contiguous */
  DG_BLK_N_H2, DG_BLK_N_H3,          /* scratch [B*N, H] bf16 */
  DG_BLK_ES0, DG_BLK_ES1, DG_BLK_ES2, DG_BLK_ES3, DG_BLK_ES4, DG_BLK_ES5, DG_BLK_ES6, DG_BLK_ES7, DG_BLK_ES8, /* scratch, edge fp32
                    * (without a live edge output only ES0 and ES5..ES8 are used) */
  DG_BLK_E_H2, DG_BLK_E_H3,          /* scratch [B*N*N, H] bf16 (only with a live edge output) */
  DG_BLK_WT,       /* scratch, 2 H D floats (transposed MLP weights) */
  DG_BLK_COUNT
};
#define DG_BLK_BB_NODE_SLOTS 28
#define DG_BLKF_EDGE_OUT 1 /* the block's edge output has a consumer (every block but the Discriminator's last, models.py:202-207) */
#define DG_BLKF_KEEP 2     /* fwd: also write E and Z4 (the caller keeps X1..X3, Y3, A16, E, Z4 for dg_block_bwd); bwd: they are valid */
#define DG_BLKF_STATS 4    /* fwd: write STAT_M / STAT_INV; bwd: G / STAT_M / STAT_INV hold the forward's values */
/* x_out, y_out = Encoder_Block.forward(x, y). */
int dg_block_fwd(void* const* io, const float* const* params, int B, int N, int D, int H, int heads, int flags, float eps,
                 void* workspace, long long workspace_bytes, void* stream);
/* dx, dy and the parameter gradients of one block from d x_out / d y_out (what autograd runs for layers.py:174-193).  Without
 * DG_BLKF_KEEP the forward intermediates are recomputed from (x, y) into their io slots first.  `grads`: DG_BLOCK_PARAMS device
 * pointers, ZEROED by the caller, accumulated into (+=); NULL = no parameter gradients (dgrad-only pass: the gradient penalty's
 * input gradient, the Generator step's pass through the Discriminator).  Entries of parameters without a consumer (out_e, ln4,
 * mlp2, ln6 when the edge output is not live) may be NULL and are left untouched. */
int dg_block_bwd(void* const* io, const float* const* params, float* const* grads, int B, int N, int D, int H, int heads, int flags,
                 float eps, void* workspace, long long workspace_bytes, void* stream);
/* Second-order pass of one block for the gradient penalty's double backward (loss.py:32-39 inside d_loss.backward()): the
 * gradient of <ux, dx> + <uy, dy>, (dx, dy) = dg_block_bwd(x, y, dxo, dyo), with respect to x, y, dxo, dyo and the parameters --
 * hand-sequenced (no autograd graph): recompute the forward and the first-order backward, walk the backward program in reverse
 * with the second-order kernels, then one first-order backward from the cotangents injected at the forward intermediates.
 * DG_BLKF_KEEP: X1, Q, K, V, Y3, E, Z4 hold the forward's values (else recomputed into their slots); G, STAT_M, STAT_INV, ON, X3 are
 * always scratch here (the softmax is retaken from fp32 scores).  UX, UY, DXO are required (zeros where there is no cotangent);
 * DYO NULL = the edge output is not live.  `grads`: as for dg_block_bwd (zeroed by the caller, accumulated into), required. */
int dg_block_bwd_bwd(void* const* io, const float* const* params, float* const* grads, int B, int N, int D, int H, int heads,
                     int flags, float eps, void* workspace, long long workspace_bytes, void* stream);
/* TransformerEncoder.forward (layers.py:221-234) without a graph: `depth` blocks, params = depth x DG_BLOCK_PARAMS pointers.
 * `scratch`: a DG_BLK_* table holding X1..X3, Y3, A16 (E when !last_edge_out) and, for depth > 1, ping-pong buffers in the
 * X_OUT / Y_OUT slots; x_out / y_out also serve as ping-pong buffers (y_out may be NULL only if depth <= 2 and !last_edge_out). */
int dg_encoder_fwd(const float* x, const float* y, float* x_out, float* y_out, const float* const* params, int depth,
                   void* const* scratch, int B, int N, int D, int H, int heads, int last_edge_out, float eps, void* workspace,
                   long long workspace_bytes, void* stream);
/* Launches issued by the block-level entry points so far (process-wide). */
long long dg_native_launches(void);
/* Launch probe: every launch issued by the block-level entry points whose key (the text druggen_b200/_lib.py gives the same
 * launch, e.g. "mlp_bwd_ln[R=4147200,H=384,fused,mask]") equals `key` is bracketed by CUDA events on its stream; NULL / "" stops.
 * dg_probe_read: waits for the recorded launches, returns their number and summed duration, and forgets them. */
int dg_probe_set(const char* key);
int dg_probe_read(long long* launches, double* total_ms);

/* debug / tests: dry-run trace.  With the trace on, every kernel entry point above records one line -- its name and its arguments
 * (p:<hex pointer>, i:<integer>, f:<float>) -- and returns 0 WITHOUT touching the device; the block-level entry points then list
 * their launch programs (plus memset0 / transpose / add3 lines for their own small launches), which tests/test_native_trace.py
 * pins on a box without a GPU.  dg_debug_trace_read copies the text recorded so far (NUL-terminated, at most cap - 1 bytes) and
 * clears it; returns its length. */
int dg_debug_trace(int on);
long long dg_debug_trace_read(char* buf, long long cap);
/* debug: with a device buffer of 148*64 int64 set, every chain-kernel launch (dg_mlp_*, dg_attn_edge_fwd) writes
 * per-CTA phase cycle counters [CTA][4 roles][16 phases] (tools/chain_profile.py); NULL switches it off. */
int dg_debug_chain_profile(void* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* DRUGGEN_B200_H */

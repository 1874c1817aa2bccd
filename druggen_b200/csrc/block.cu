// Block-level entry points of the C-ABI (dg_block_fwd / dg_block_bwd): ONE host call per direction of an encoder
// block (layers.py:174-193).  A call is a fixed sequence of the library's own kernels -- the same launches, with the same
// arguments, that druggen_b200/block.py issues one by one from Python in the tensor-core throughput mode -- over buffers the
// caller owns (a table of device pointers, DG_BLK_*): nothing is allocated here, nothing synchronises, and every launch goes to
// the caller's stream, so a call can be captured into a CUDA graph (dg_encoder_fwd: the launch-bound small-batch forward).
// What this removes is the per-launch host cost of the Python sequencing (~100 us per launch x ~2200 launches per GAN step:
// the whole step at 512 molecules), not device time.
//
// Launch probe: bench.py times the dominant kernel live inside its timed region.  With dg_probe_set(key) every launch issued
// from here whose key -- the same text druggen_b200/_lib.py gives that launch -- matches is bracketed by CUDA events on the
// launching stream; dg_probe_read() sums them.
#include <cmath>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace {

std::mutex g_mu;
char g_probe_key[128] = {0};
bool g_probe_on = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_probe_used, g_probe_free;
long long g_native_launches = 0;

// One launch sequence on one stream; the first rejected launch stops it (dg_last_error() keeps that launch's text).
struct Seq {
  cudaStream_t s;
  int rc = 0;
  long long n = 0;
  cudaEvent_t e1 = nullptr;
  explicit Seq(void* stream) : s((cudaStream_t)stream) {}
  // key: printf-style text of the launch, only formatted while a probe is set
  bool before(const char* fmt, ...) {
    if (rc) return false;
    ++n;
    if (!g_probe_on) return true;
    char key[128];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(key, sizeof key, fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(g_mu);
    if (std::strcmp(key, g_probe_key) != 0) return true;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!g_probe_free.empty()) {
      ev = g_probe_free.back();
      g_probe_free.pop_back();
    } else if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) {
      return true;
    }
    cudaEventRecord(ev.first, s);
    e1 = ev.second;
    g_probe_used.push_back(ev);
    return true;
  }
  void after(int r) {
    rc = r;
    if (e1) {
      cudaEventRecord(e1, s);
      e1 = nullptr;
    }
  }
  int done() {
    std::lock_guard<std::mutex> lk(g_mu);
    g_native_launches += n;
    return rc;
  }
};

#define DG_STEP(seq, call, ...)        \
  do {                                 \
    if ((seq).before(__VA_ARGS__)) (seq).after(call); \
  } while (0)

const char* kPrec = "bf16";
constexpr int kP = DG_PREC_BF16;

// parameter table: the reference's state-dict order of a block (druggen_b200/block.py BLOCK_PARAM_NAMES)
enum {
  LN1_W, LN1_B, Q_W, Q_B, K_W, K_B, V_W, V_B, E_W, E_B, OE_W, OE_B, ON_W, ON_B, LN3_W, LN3_B, LN4_W, LN4_B,
  FC1_W, FC1_B, FC2_W, FC2_B, FC1E_W, FC1E_B, FC2E_W, FC2E_B, LN5_W, LN5_B, LN6_W, LN6_B, kNumParams
};
static_assert(kNumParams == DG_BLOCK_PARAMS, "parameter table");

int shape_check(const char* who, int B, int N, int D, int H, int heads, long long ws_bytes, const void* ws) {
  if (B <= 0 || N < 4 || N > 212 || B > 65535) return dg::fail("%s: needs 4 <= N <= 212 and 0 < B <= 65535, got B=%d N=%d", who, B, N);
  if (D != 128 || H % 128 || H < 128 || H > 384) return dg::fail("%s: needs D == 128 and H in {128,256,384}, got D=%d H=%d", who, D, H);
  if (heads <= 0 || D % heads) return dg::fail("%s: heads must divide D, got %d", who, heads);
  if ((long long)B * N * N >= (1ll << 31) || (long long)B * N * 128 >= (1ll << 31)) return dg::fail("%s: B*N*N and B*N*128 must be < 2^31 (split the batch)", who);
  if (ws_bytes < (long long)2 * (H / 128) * 32768 || ws_bytes < 65536) return dg::fail("%s: workspace too small (%lld bytes)", who, ws_bytes);
  if (reinterpret_cast<uintptr_t>(ws) & 127) return dg::fail("%s: workspace must be 128-byte aligned", who);
  return 0;
}

template <class T = float>
T* at(void* const* io, int i) { return reinterpret_cast<T*>(io[i]); }

int need(const char* who, void* const* io, std::initializer_list<int> idx) {
  for (int i : idx)
    if (io[i] == nullptr) return dg::fail("%s: buffer %d of the DG_BLK_* table is required for this call", who, i);
  return 0;
}

int zero(Seq& q, void* p, long long bytes) {
  if (q.rc) return q.rc;
  if (dg::trace_on()) return dg::trace_call("memset0", (const void*)p, bytes);
  if (cudaMemsetAsync(p, 0, (size_t)bytes, q.s) != cudaSuccess) q.rc = dg::fail("dg_block: memset failed");
  return q.rc;
}

// x1 = LN1(x); q, k, v = their projections (layers.py:185,111-113)
void node_prologue(Seq& q, void* const* io, const float* const* P, long long BN, int D, float eps) {
  DG_STEP(q, dg_add_ln_fwd(at(io, DG_BLK_X), nullptr, P[LN1_W], P[LN1_B], at(io, DG_BLK_X1), BN, D, eps, q.s), "add_ln_fwd");
  const int w[3] = {Q_W, K_W, V_W}, o[3] = {DG_BLK_Q, DG_BLK_K, DG_BLK_V};
  for (int i = 0; i < 3; ++i)
    DG_STEP(q, dg_rows_gemm(io[DG_BLK_X1], P[w[i]], 1, P[w[i] + 1], 0, nullptr, nullptr, io[o[i]], BN, D, D, kP, 0, q.s),
            "rows_gemm[R=%lld,K=%d,N=%d,%s]", BN, D, D, kPrec);
}

// on = out_n(g); x3 = LN3(x1 + on) (layers.py:135,187,189)
void node_epilogue(Seq& q, void* const* io, const float* const* P, long long BN, int D, float eps) {
  DG_STEP(q, dg_rows_gemm(io[DG_BLK_G], P[ON_W], 1, P[ON_B], 0, nullptr, nullptr, io[DG_BLK_ON], BN, D, D, kP, 0, q.s),
          "rows_gemm[R=%lld,K=%d,N=%d,%s]", BN, D, D, kPrec);
  DG_STEP(q, dg_add_ln_fwd(at(io, DG_BLK_X1), at(io, DG_BLK_ON), P[LN3_W], P[LN3_B], at(io, DG_BLK_X3), BN, D, eps, q.s), "add_ln_fwd");
}

// Backward of LN(xin + fc2(relu(fc1(xin)))) given d(out): the two fused chains with the weight gradients between them.
//   dz, hbuf (bf16 [rows,H]: h, then dh), mask: scratch;  returns d(xin) in dxin.  grads == nullptr: dgrad only.
void mlp_backward(Seq& q, const float* xin, const float* dout, const float* const* P, int fc1, int ln, float* const* grads, float* dz,
                  void* hbuf, void* mask, float* dxin, long long rows, int D, int H, float eps, void* ws, long long ws_bytes) {
  const bool wp = grads != nullptr;
  const char* tag_h = wp ? "" : ",no h";
  DG_STEP(q, dg_mlp_bwd_ln(xin, dout, P[fc1], P[fc1 + 1], P[fc1 + 2], P[fc1 + 3], P[ln], dz, wp ? hbuf : nullptr, mask,
                           wp ? grads[ln] : nullptr, wp ? grads[ln + 1] : nullptr, rows, D, H, eps, ws, ws_bytes, q.s),
          "mlp_bwd_ln[R=%lld,H=%d,fused%s,mask]", rows, H, tag_h);
  if (wp)      // dW2 = dz^T h, db2 = colsum(dz)
    DG_STEP(q, dg_gemm_tn(dz, hbuf, grads[fc1 + 2], grads[fc1 + 3], rows, D, H, kP, DG_OUT_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,b16]", rows, D, H, kPrec);
  DG_STEP(q, dg_mlp_bwd_dgrad(dz, nullptr, mask, P[fc1], P[fc1 + 2], dxin, wp ? hbuf : nullptr, rows, D, H, ws, ws_bytes, q.s),
          "mlp_bwd_dgrad[R=%lld,H=%d,fused,mask%s]", rows, H, wp ? "" : ",no dh");
  if (wp)      // dW1 = dh^T xin, db1 = colsum(dh)
    DG_STEP(q, dg_gemm_tn(hbuf, xin, grads[fc1], grads[fc1 + 1], rows, H, D, kP, DG_A_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,a16]", rows, H, D, kPrec);
}

}  // namespace

extern "C" long long dg_native_launches(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_native_launches;
}

extern "C" int dg_probe_set(const char* key) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_probe_on = key != nullptr && key[0] != 0;
  std::snprintf(g_probe_key, sizeof g_probe_key, "%s", g_probe_on ? key : "");
  return 0;
}

extern "C" int dg_probe_read(long long* launches, double* total_ms) {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> used;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    used.swap(g_probe_used);
  }
  double ms = 0.0;
  long long n = 0;
  for (auto& ev : used) {
    float t = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&t, ev.first, ev.second) == cudaSuccess) {
      ms += t;
      ++n;
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& ev : used) g_probe_free.push_back(ev);
  }
  if (launches) *launches = n;
  if (total_ms) *total_ms = ms;
  return 0;
}

extern "C" int dg_block_fwd(void* const* io, const float* const* P, int B, int N, int D, int H, int heads, int flags, float eps,
                            void* ws, long long ws_bytes, void* stream) {
  if (shape_check("dg_block_fwd", B, N, D, H, heads, ws_bytes, ws)) return 1;
  const bool edge_out = flags & DG_BLKF_EDGE_OUT, keep = flags & DG_BLKF_KEEP, stats = flags & DG_BLKF_STATS;
  if (need("dg_block_fwd", io, {DG_BLK_X, DG_BLK_Y, DG_BLK_X_OUT, DG_BLK_X1, DG_BLK_Q, DG_BLK_K, DG_BLK_V, DG_BLK_G, DG_BLK_ON, DG_BLK_X3})) return 1;
  if (edge_out && need("dg_block_fwd", io, {DG_BLK_Y_OUT, DG_BLK_Y3, DG_BLK_A16})) return 1;
  if ((!edge_out || keep) && need("dg_block_fwd", io, {DG_BLK_E})) return 1;
  if (!edge_out && need("dg_block_fwd", io, {DG_BLK_Y3})) return 1;
  if (keep && (!edge_out || need("dg_block_fwd", io, {DG_BLK_Z4}))) return dg::fail("dg_block_fwd: DG_BLKF_KEEP needs DG_BLKF_EDGE_OUT and the E / Z4 buffers");
  if (stats && need("dg_block_fwd", io, {DG_BLK_STAT_M, DG_BLK_STAT_INV})) return 1;
  const long long BN = (long long)B * N, R = BN * N;
  const float c = 1.0f / std::sqrt((float)(D / heads));                                        // layers.py:124
  Seq q(stream);
  node_prologue(q, io, P, BN, D, eps);
  float* sm = stats ? at(io, DG_BLK_STAT_M) : nullptr;
  float* si = stats ? at(io, DG_BLK_STAT_INV) : nullptr;
  if (edge_out) {
    // E-projection, modulation, out_e projection, residual, LN4 in one tcgen05 kernel; the scores leave once, as bf16
    DG_STEP(q, dg_attn_edge_fwd(at(io, DG_BLK_Y), at(io, DG_BLK_Q), at(io, DG_BLK_K), P[E_W], P[E_B], P[OE_W], P[OE_B], P[LN4_W], P[LN4_B], c,
                                at(io, DG_BLK_Y3), io[DG_BLK_A16], keep ? at(io, DG_BLK_E) : nullptr, keep ? at(io, DG_BLK_Z4) : nullptr, B, N,
                                D, eps, ws, ws_bytes, q.s),
            "attn_edge_fwd[fused+a16%s]", keep ? "+e+z" : "");
    DG_STEP(q, dg_softmax_agg16_fwd(io[DG_BLK_A16], at(io, DG_BLK_V), at(io, DG_BLK_G), sm, si, B, N, D, q.s), "softmax_agg16_fwd");
  } else {
    // the block whose edge output nobody reads (the last Discriminator block, models.py:202-207): E, then softmax-aggregate only
    DG_STEP(q, dg_rows_gemm(io[DG_BLK_Y], P[E_W], 1, P[E_B], 0, nullptr, nullptr, io[DG_BLK_E], R, D, D, kP, 0, q.s),
            "rows_gemm[R=%lld,K=%d,N=%d,%s]", R, D, D, kPrec);
    // (the scores are stored too, into the Y3 slot, although nothing reads them: the score-storing kernel is the one block.py runs
    //  here in every precision mode, and the two softmax kernels differ in the last bits)
    DG_STEP(q, dg_attn_scores_fwd(at(io, DG_BLK_Q), at(io, DG_BLK_K), at(io, DG_BLK_V), at(io, DG_BLK_E), c, at(io, DG_BLK_Y3), at(io, DG_BLK_G),
                                  sm, si, B, N, D, q.s),
            "attn_scores_fwd[fused]");
  }
  node_epilogue(q, io, P, BN, D, eps);
  DG_STEP(q, dg_mlp_fwd(at(io, DG_BLK_X3), P[FC1_W], P[FC1_B], P[FC2_W], P[FC2_B], P[LN5_W], P[LN5_B], at(io, DG_BLK_X_OUT), BN, D, H, eps,
                        ws, ws_bytes, q.s),
          "mlp_fwd[R=%lld,H=%d,fused]", BN, H);
  if (edge_out)
    DG_STEP(q, dg_mlp_fwd(at(io, DG_BLK_Y3), P[FC1E_W], P[FC1E_B], P[FC2E_W], P[FC2E_B], P[LN6_W], P[LN6_B], at(io, DG_BLK_Y_OUT), R, D, H,
                          eps, ws, ws_bytes, q.s),
            "mlp_fwd[R=%lld,H=%d,fused]", R, H);
  return q.done();
}

extern "C" int dg_encoder_fwd(const float* x, const float* y, float* x_out, float* y_out, const float* const* params, int depth,
                              void* const* scratch, int B, int N, int D, int H, int heads, int last_edge_out, float eps, void* ws,
                              long long ws_bytes, void* stream) {
  if (depth <= 0) return dg::fail("dg_encoder_fwd: depth must be > 0");
  if (need("dg_encoder_fwd", scratch, {DG_BLK_X1, DG_BLK_Q, DG_BLK_K, DG_BLK_V, DG_BLK_G, DG_BLK_ON, DG_BLK_X3, DG_BLK_Y3, DG_BLK_A16})) return 1;
  if (depth > 1 && (scratch[DG_BLK_X_OUT] == nullptr || scratch[DG_BLK_Y_OUT] == nullptr))
    return dg::fail("dg_encoder_fwd: depth > 1 needs the ping-pong buffers in the X_OUT / Y_OUT slots of `scratch`");
  if (!last_edge_out && scratch[DG_BLK_E] == nullptr) return dg::fail("dg_encoder_fwd: the block without an edge output needs the E buffer");
  // layer l reads (cx, cy) and writes the caller's outputs (last layer) or a ping-pong buffer: layer l -> buffer (depth-1-l) & 1,
  // so that the last layer's input is never the buffer it writes; buffer 1 = the scratch X_OUT / Y_OUT slots, buffer 0 = the
  // caller's outputs themselves (they are dead until the last layer writes them)
  const float *cx = x, *cy = y;
  for (int l = 0; l < depth; ++l) {
    void* io[DG_BLK_COUNT];
    for (int i = 0; i < DG_BLK_COUNT; ++i) io[i] = scratch[i];
    const bool last = l == depth - 1;
    const bool to_scratch = !last && ((depth - 1 - l) & 1);
    const bool eo = !last || last_edge_out;
    io[DG_BLK_X] = const_cast<float*>(cx);
    io[DG_BLK_Y] = const_cast<float*>(cy);
    io[DG_BLK_X_OUT] = to_scratch ? scratch[DG_BLK_X_OUT] : x_out;
    io[DG_BLK_Y_OUT] = to_scratch ? scratch[DG_BLK_Y_OUT] : y_out;
    if (eo && io[DG_BLK_Y_OUT] == nullptr) return dg::fail("dg_encoder_fwd: y_out is required when the last block has an edge output");
    if (int rc = dg_block_fwd(io, params + (long long)l * DG_BLOCK_PARAMS, B, N, D, H, heads, eo ? DG_BLKF_EDGE_OUT : 0, eps, ws, ws_bytes, stream))
      return rc;
    cx = at(io, DG_BLK_X_OUT);
    cy = at(io, DG_BLK_Y_OUT);
  }
  return 0;
}

extern "C" int dg_block_bwd(void* const* io, const float* const* P, float* const* grads, int B, int N, int D, int H, int heads, int flags,
                            float eps, void* ws, long long ws_bytes, void* stream) {
  if (shape_check("dg_block_bwd", B, N, D, H, heads, ws_bytes, ws)) return 1;
  const bool edge_out = flags & DG_BLKF_EDGE_OUT, kept = flags & DG_BLKF_KEEP, have_stats = flags & DG_BLKF_STATS;
  const bool live = edge_out && io[DG_BLK_DYO] != nullptr;
  if (kept && !(live && have_stats)) return dg::fail("dg_block_bwd: kept intermediates need a live edge output and the forward's statistics");
  if (need("dg_block_bwd", io, {DG_BLK_X, DG_BLK_Y, DG_BLK_X1, DG_BLK_Q, DG_BLK_K, DG_BLK_V, DG_BLK_G, DG_BLK_ON, DG_BLK_X3, DG_BLK_STAT_M,
                                DG_BLK_STAT_INV, DG_BLK_E, DG_BLK_DX, DG_BLK_DY, DG_BLK_N_DZ, DG_BLK_N_MASK, DG_BLK_N_DX3, DG_BLK_N_DZ3,
                                DG_BLK_N_DG, DG_BLK_N_DQ, DG_BLK_N_DK, DG_BLK_N_DV, DG_BLK_N_T0, DG_BLK_N_T1, DG_BLK_E_H, DG_BLK_SCRATCH}))
    return 1;
  if (live && need("dg_block_bwd", io, {DG_BLK_Y3, DG_BLK_A16, DG_BLK_Z4, DG_BLK_E_A, DG_BLK_E_B, DG_BLK_E_MASK})) return 1;
  if (!live && need("dg_block_bwd", io, {DG_BLK_Y3})) return 1;      // (scratch for the scores of the recomputed softmax)
  const bool wp = grads != nullptr;
  if (wp && need("dg_block_bwd", io, {DG_BLK_N_H})) return 1;
  if (wp)
    for (int i = 0; i < kNumParams; ++i) {
      const bool edge_only = i == OE_W || i == OE_B || i == LN4_W || i == LN4_B || (i >= FC1E_W && i <= FC2E_B) || i == LN6_W || i == LN6_B;
      if (grads[i] == nullptr && (live || !edge_only)) return dg::fail("dg_block_bwd: gradient buffer %d is missing", i);
    }
  const long long BN = (long long)B * N, R = BN * N;
  const float c = 1.0f / std::sqrt((float)(D / heads));
  float* scratch = at(io, DG_BLK_SCRATCH);            // 2 D floats: LayerNorm affine gradients nobody reads (dgrad-only passes)
  Seq q(stream);
  // ---- recompute the forward intermediates from the block inputs, unless the forward kept them
  if (!kept) {
    node_prologue(q, io, P, BN, D, eps);
    if (live) {
      DG_STEP(q, dg_attn_edge_fwd(at(io, DG_BLK_Y), at(io, DG_BLK_Q), at(io, DG_BLK_K), P[E_W], P[E_B], P[OE_W], P[OE_B], P[LN4_W], P[LN4_B], c,
                                  at(io, DG_BLK_Y3), io[DG_BLK_A16], at(io, DG_BLK_E), at(io, DG_BLK_Z4), B, N, D, eps, ws, ws_bytes, q.s),
              "attn_edge_fwd[fused+a16+e+z]");
      if (!have_stats)        // (the checkpointed forward normally hands its statistics over: node-sized)
        DG_STEP(q, dg_softmax_agg16_fwd(io[DG_BLK_A16], at(io, DG_BLK_V), at(io, DG_BLK_G), at(io, DG_BLK_STAT_M), at(io, DG_BLK_STAT_INV), B,
                                        N, D, q.s),
                "softmax_agg16_fwd");
    } else {
      DG_STEP(q, dg_rows_gemm(io[DG_BLK_Y], P[E_W], 1, P[E_B], 0, nullptr, nullptr, io[DG_BLK_E], R, D, D, kP, 0, q.s),
              "rows_gemm[R=%lld,K=%d,N=%d,%s]", R, D, D, kPrec);
      DG_STEP(q, dg_attn_scores_fwd(at(io, DG_BLK_Q), at(io, DG_BLK_K), at(io, DG_BLK_V), at(io, DG_BLK_E), c, at(io, DG_BLK_Y3), at(io, DG_BLK_G),
                                    at(io, DG_BLK_STAT_M), at(io, DG_BLK_STAT_INV), B, N, D, q.s),
              "attn_scores_fwd[fused]");
    }
    node_epilogue(q, io, P, BN, D, eps);
  }
  // ---- node stream: MLP + LN5, LN3, out_n
  const float* dxo = at(io, DG_BLK_DXO);
  if (dxo == nullptr) {
    zero(q, io[DG_BLK_N_T0], BN * D * (long long)sizeof(float));
    dxo = at(io, DG_BLK_N_T0);
  }
  mlp_backward(q, at(io, DG_BLK_X3), dxo, P, FC1_W, LN5_W, wp ? grads : nullptr, at(io, DG_BLK_N_DZ), io[DG_BLK_N_H], io[DG_BLK_N_MASK],
               at(io, DG_BLK_N_DX3), BN, D, H, eps, ws, ws_bytes);
  // dz3: the gradient of both x1 (residual) and out_n(g)
  DG_STEP(q, dg_add_ln_bwd(at(io, DG_BLK_N_DX3), at(io, DG_BLK_X1), at(io, DG_BLK_ON), P[LN3_W], at(io, DG_BLK_N_DZ3), wp ? grads[LN3_W] : scratch,
                           wp ? grads[LN3_B] : scratch + D, BN, D, eps, 0, q.s),
          "add_ln_bwd");
  if (wp)
    DG_STEP(q, dg_gemm_tn(io[DG_BLK_N_DZ3], io[DG_BLK_G], grads[ON_W], grads[ON_B], BN, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", BN, D, D, kPrec);
  DG_STEP(q, dg_rows_gemm(io[DG_BLK_N_DZ3], P[ON_W], 0, nullptr, 0, nullptr, nullptr, io[DG_BLK_N_DG], BN, D, D, kP, 0, q.s),
          "rows_gemm[R=%lld,K=%d,N=%d,%s]", BN, D, D, kPrec);
  // ---- edge stream: MLP2 + LN6, LN4, out_e
  const void* da = nullptr;
  float* dz4 = nullptr;
  if (live) {
    float* dy3 = at(io, DG_BLK_E_B);
    mlp_backward(q, at(io, DG_BLK_Y3), at(io, DG_BLK_DYO), P, FC1E_W, LN6_W, wp ? grads : nullptr, at(io, DG_BLK_E_A), io[DG_BLK_E_H],
                 io[DG_BLK_E_MASK], dy3, R, D, H, eps, ws, ws_bytes);
    dz4 = at(io, DG_BLK_E_A);                         // (the MLP's dz is dead: its buffer takes the gradient of y + out_e(A))
    DG_STEP(q, dg_add_ln_bwd(dy3, at(io, DG_BLK_Z4), nullptr, P[LN4_W], dz4, wp ? grads[LN4_W] : scratch, wp ? grads[LN4_B] : scratch + D, R, D,
                             eps, 0, q.s),
            "add_ln_bwd");
    if (wp)
      DG_STEP(q, dg_gemm_tn(dz4, io[DG_BLK_A16], grads[OE_W], grads[OE_B], R, D, D, kP, DG_OUT_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,b16]", R, D, D, kPrec);
    // the out_e path's gradient of the scores leaves as bf16 (it is only ever added to the softmax term inside the next kernel)
    DG_STEP(q, dg_rows_gemm(dz4, P[OE_W], 0, nullptr, 0, nullptr, nullptr, dy3, R, D, D, kP, DG_OUT_BF16, q.s),
            "rows_gemm[R=%lld,K=%d,N=%d,%s,o16]", R, D, D, kPrec);
    da = dy3;
  }
  // ---- attention: softmax-aggregate + modulation backward (dE as bf16: only ever a contraction operand)
  for (int slot : {DG_BLK_N_DQ, DG_BLK_N_DK, DG_BLK_N_DV}) zero(q, io[slot], BN * D * (long long)sizeof(float));
  void* de = io[DG_BLK_E_H];
  DG_STEP(q, dg_attn_scores_bwd(at(io, DG_BLK_N_DG), (const float*)da, at(io, DG_BLK_Q), at(io, DG_BLK_K), at(io, DG_BLK_V), at(io, DG_BLK_E), c,
                                at(io, DG_BLK_STAT_M), at(io, DG_BLK_STAT_INV), at(io, DG_BLK_G), de, at(io, DG_BLK_N_DQ), at(io, DG_BLK_N_DK),
                                at(io, DG_BLK_N_DV), B, N, D, 1 | (live ? 2 : 0) | (da ? 4 : 0), q.s),
          "attn_scores_bwd[fused,de16%s]", da ? ",da16" : "");
  if (wp)
    DG_STEP(q, dg_gemm_tn(de, io[DG_BLK_Y], grads[E_W], grads[E_B], R, D, D, kP, DG_A_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,a16]", R, D, D, kPrec);
  DG_STEP(q, dg_rows_gemm(de, P[E_W], 0, nullptr, 0, nullptr, dz4, io[DG_BLK_DY], R, D, D, kP, DG_A_BF16, q.s),
          "rows_gemm[R=%lld,K=%d,N=%d,%s%s,a16]", R, D, D, kPrec, dz4 ? "+resid" : "");
  // ---- q / k / v projections and LN1
  const float* dx1 = at(io, DG_BLK_N_DZ3);
  const int w[3] = {Q_W, K_W, V_W}, dt[3] = {DG_BLK_N_DQ, DG_BLK_N_DK, DG_BLK_N_DV}, out[3] = {DG_BLK_N_T0, DG_BLK_N_T1, DG_BLK_N_T0};
  for (int i = 0; i < 3; ++i) {
    if (wp)
      DG_STEP(q, dg_gemm_tn(io[dt[i]], io[DG_BLK_X1], grads[w[i]], grads[w[i] + 1], BN, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", BN, D, D, kPrec);
    DG_STEP(q, dg_rows_gemm(io[dt[i]], P[w[i]], 0, nullptr, 0, nullptr, dx1, io[out[i]], BN, D, D, kP, 0, q.s),
            "rows_gemm[R=%lld,K=%d,N=%d,%s+resid]", BN, D, D, kPrec);
    dx1 = at(io, out[i]);
  }
  DG_STEP(q, dg_add_ln_bwd(dx1, at(io, DG_BLK_X), nullptr, P[LN1_W], at(io, DG_BLK_DX), wp ? grads[LN1_W] : scratch, wp ? grads[LN1_B] : scratch + D,
                           BN, D, eps, 0, q.s),
          "add_ln_bwd");
  return q.done();
}

// ---- second-order pass ----------------------------------------------------------------------------------------------------
namespace {

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  // out[c, r] = in[r, c]; the MLP weights (<= 384 x 128): a few dozen KB, one pass
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
    const int r = i / cols, c = i - r * cols;
    out[(long long)c * rows + r] = in[i];
  }
}

// a[i] += b[i] over three equally sized pairs in one launch (c[q], c[k], c[v] += their second contributions)
__global__ void add3_kernel(float4* a0, const float4* b0, float4* a1, const float4* b1, float4* a2, const float4* b2, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 x = a0[i], y = b0[i];
    a0[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    x = a1[i]; y = b1[i];
    a1[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    x = a2[i]; y = b2[i];
    a2[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}

// node-sized scratch slots of dg_block_bwd_bwd inside DG_BLK_N_ARENA
enum {
  nT5, nM, nDX3, nDZ3, nDG, nDV, nDQ, nDK, nP0, nP1, nCDX1, nCDQ, nCDK, nCDV, nCQ, nCK, nCV, nCDG, nCDZ3, nCDX3, nCZ3, nCT5, nCMN,
  nCX3, nCG, nDQ2, nDK2, nDV2, kNodeSlots
};
static_assert(kNodeSlots == DG_BLK_BB_NODE_SLOTS, "node arena");

// forward m = xin + fc2(h) and the first-order backward (t = LN^T dout, dh, dxin) of LN(xin + mlp(xin)), intermediates kept
void mlp_recompute(Seq& q, const float* xin, const float* dout, const float* const* P, int fc1, int ln, float* t, void* h, void* mask,
                   float* m, float* dxin, void* dh, long long rows, int D, int H, float eps, void* ws, long long wsb) {
  DG_STEP(q, dg_mlp_bwd_ln(xin, dout, P[fc1], P[fc1 + 1], P[fc1 + 2], P[fc1 + 3], P[ln], t, h, mask, nullptr, nullptr, rows, D, H, eps, ws, wsb, q.s),
          "mlp_bwd_ln[R=%lld,H=%d,fused,mask]", rows, H);
  DG_STEP(q, dg_rows_gemm(h, P[fc1 + 2], 1, P[fc1 + 3], 0, nullptr, xin, m, rows, H, D, kP, DG_A_BF16, q.s),
          "rows_gemm[R=%lld,K=%d,N=%d,%s+resid,a16]", rows, H, D, kPrec);
  DG_STEP(q, dg_mlp_bwd_dgrad(t, nullptr, mask, P[fc1], P[fc1 + 2], dxin, dh, rows, D, H, ws, wsb, q.s), "mlp_bwd_dgrad[R=%lld,H=%d,fused,mask]", rows, H);
}

// reverse of dxin = t + ((t W2) * M) W1 given u = c[dxin]: c[t] into c_t; c[W2] += t^T tM, c[W1] += dh^T u
void mlp_second(Seq& q, const float* u, const float* t, const void* mask, const void* dh, const float* const* P, int fc1, float* const* grads,
                float* c_t, void* tm, float* wt, long long rows, int D, int H, void* ws, long long wsb) {
  if (!q.rc && dg::trace_on()) {
    dg::trace_call("transpose", (const void*)P[fc1 + 2], (const void*)wt, D, H);
    dg::trace_call("transpose", (const void*)P[fc1], (const void*)(wt + (long long)H * D), H, D);
  } else if (!q.rc) {      // the dgrad chain with the two weights transposed into each other's role
    transpose_kernel<<<96, 256, 0, q.s>>>(P[fc1 + 2], wt, D, H);                  // W2 [D,H] -> [H,D]
    transpose_kernel<<<96, 256, 0, q.s>>>(P[fc1], wt + (long long)H * D, H, D);   // W1 [H,D] -> [D,H]
  }
  DG_STEP(q, dg_mlp_bwd_dgrad(u, nullptr, mask, wt, wt + (long long)H * D, c_t, tm, rows, D, H, ws, wsb, q.s),
          "mlp_bwd_dgrad[R=%lld,H=%d,fused,mask]", rows, H);
  DG_STEP(q, dg_gemm_tn(t, tm, grads[fc1 + 2], nullptr, rows, D, H, kP, DG_OUT_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,b16]", rows, D, H, kPrec);
  DG_STEP(q, dg_gemm_tn(dh, u, grads[fc1], nullptr, rows, H, D, kP, DG_A_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,a16]", rows, H, D, kPrec);
}

// reverse of m = xin + fc2(relu(fc1(xin))) given c[m]: c[xin] into c_xin; the four parameter cotangents accumulate
void mlp_first(Seq& q, const float* c_m, const void* h, const void* mask, const float* xin, const float* const* P, int fc1,
               float* const* grads, float* c_xin, void* ch, long long rows, int D, int H, void* ws, long long wsb) {
  DG_STEP(q, dg_mlp_bwd_dgrad(c_m, nullptr, mask, P[fc1], P[fc1 + 2], c_xin, ch, rows, D, H, ws, wsb, q.s),
          "mlp_bwd_dgrad[R=%lld,H=%d,fused,mask]", rows, H);
  DG_STEP(q, dg_gemm_tn(c_m, h, grads[fc1 + 2], grads[fc1 + 3], rows, D, H, kP, DG_OUT_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,b16]", rows, D, H, kPrec);
  DG_STEP(q, dg_gemm_tn(ch, xin, grads[fc1], grads[fc1 + 1], rows, H, D, kP, DG_A_BF16, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s,a16]", rows, H, D, kPrec);
}

}  // namespace

extern "C" int dg_block_bwd_bwd(void* const* io, const float* const* P, float* const* grads, int B, int N, int D, int H, int heads,
                                int flags, float eps, void* ws, long long wsb, void* stream) {
  const char* who = "dg_block_bwd_bwd";
  if (shape_check(who, B, N, D, H, heads, wsb, ws)) return 1;
  const bool edge_out = flags & DG_BLKF_EDGE_OUT, kept = flags & DG_BLKF_KEEP;
  const bool live = edge_out && io[DG_BLK_DYO] != nullptr;
  if (kept && !live) return dg::fail("%s: kept intermediates need a live edge output", who);
  if (grads == nullptr) return dg::fail("%s: the parameter cotangent table is required", who);
  if (need(who, io, {DG_BLK_X, DG_BLK_Y, DG_BLK_DXO, DG_BLK_UX, DG_BLK_UY, DG_BLK_X1, DG_BLK_Q, DG_BLK_K, DG_BLK_V, DG_BLK_G, DG_BLK_STAT_M,
                     DG_BLK_STAT_INV, DG_BLK_ON, DG_BLK_X3, DG_BLK_E, DG_BLK_C_X, DG_BLK_C_Y, DG_BLK_C_DXO, DG_BLK_N_ARENA, DG_BLK_N_H,
                     DG_BLK_N_H2, DG_BLK_N_H3, DG_BLK_N_MASK, DG_BLK_ES0, DG_BLK_ES5, DG_BLK_ES6, DG_BLK_ES7, DG_BLK_ES8, DG_BLK_WT,
                     DG_BLK_SCRATCH}))
    return 1;
  if (live && need(who, io, {DG_BLK_Y3, DG_BLK_Z4, DG_BLK_C_DYO, DG_BLK_ES1, DG_BLK_ES2, DG_BLK_ES3, DG_BLK_ES4, DG_BLK_E_H, DG_BLK_E_H2,
                             DG_BLK_E_H3, DG_BLK_E_MASK}))
    return 1;
  for (int i = 0; i < kNumParams; ++i) {
    const bool edge_only = i == OE_W || i == OE_B || i == LN4_W || i == LN4_B || (i >= FC1E_W && i <= FC2E_B) || i == LN6_W || i == LN6_B;
    if (i == LN5_B || i == LN6_B) continue;            // (the backward program does not depend on the output LayerNorms' shifts)
    if (grads[i] == nullptr && (live || !edge_only)) return dg::fail("%s: cotangent buffer %d is missing", who, i);
  }
  const long long BN = (long long)B * N, R = BN * N, nb = BN * D * (long long)sizeof(float);
  const float c = 1.0f / std::sqrt((float)(D / heads));
  float* scratch = at(io, DG_BLK_SCRATCH);
  float* arena = at(io, DG_BLK_N_ARENA);
  auto ns = [&](int slot) { return arena + (long long)slot * BN * D; };
  const float *x = at(io, DG_BLK_X), *y = at(io, DG_BLK_Y), *dxo = at(io, DG_BLK_DXO), *dyo = at(io, DG_BLK_DYO), *ux = at(io, DG_BLK_UX),
              *uy = at(io, DG_BLK_UY);
  float *x1 = at(io, DG_BLK_X1), *qq = at(io, DG_BLK_Q), *kk = at(io, DG_BLK_K), *vv = at(io, DG_BLK_V), *g = at(io, DG_BLK_G),
        *sm = at(io, DG_BLK_STAT_M), *si = at(io, DG_BLK_STAT_INV), *on = at(io, DG_BLK_ON), *x3 = at(io, DG_BLK_X3), *e = at(io, DG_BLK_E),
        *y3 = at(io, DG_BLK_Y3), *z4 = at(io, DG_BLK_Z4);
  float *s0 = at(io, DG_BLK_ES0), *s1 = at(io, DG_BLK_ES1), *s2 = at(io, DG_BLK_ES2), *s3 = at(io, DG_BLK_ES3), *s4 = at(io, DG_BLK_ES4),
        *s5 = at(io, DG_BLK_ES5), *s6 = at(io, DG_BLK_ES6), *s7 = at(io, DG_BLK_ES7), *s8 = at(io, DG_BLK_ES8);
  float* wt = at(io, DG_BLK_WT);
  const int w3[3] = {Q_W, K_W, V_W};
  Seq q(stream);
  // ---- 1a. forward recompute (the node projections and the edge chain's outputs may come from the forward)
  if (!kept) {
    node_prologue(q, io, P, BN, D, eps);
    if (live)
      DG_STEP(q, dg_attn_edge_fwd(y, qq, kk, P[E_W], P[E_B], P[OE_W], P[OE_B], P[LN4_W], P[LN4_B], c, y3, nullptr, e, z4, B, N, D, eps, ws, wsb, q.s),
              "attn_edge_fwd[fused+e+z]");
    else
      DG_STEP(q, dg_rows_gemm(y, P[E_W], 1, P[E_B], 0, nullptr, nullptr, e, R, D, D, kP, 0, q.s), "rows_gemm[R=%lld,K=%d,N=%d,%s]", R, D, D, kPrec);
  }
  float* a4 = s0;      // fp32 scores: the second-order kernels differentiate the softmax of THESE
  DG_STEP(q, dg_attn_scores_fwd(qq, kk, vv, e, c, a4, g, sm, si, B, N, D, q.s), "attn_scores_fwd[fused]");
  node_epilogue(q, io, P, BN, D, eps);
  // ---- 1b. first-order backward recompute (intermediates kept)
  mlp_recompute(q, x3, dxo, P, FC1_W, LN5_W, ns(nT5), io[DG_BLK_N_H], io[DG_BLK_N_MASK], ns(nM), ns(nDX3), io[DG_BLK_N_H2], BN, D, H, eps, ws, wsb);
  DG_STEP(q, dg_add_ln_bwd(ns(nDX3), x1, on, P[LN3_W], ns(nDZ3), scratch, scratch + D, BN, D, eps, 0, q.s), "add_ln_bwd");
  DG_STEP(q, dg_rows_gemm(ns(nDZ3), P[ON_W], 0, nullptr, 0, nullptr, nullptr, ns(nDG), BN, D, D, kP, 0, q.s), "rows_gemm[R=%lld,K=%d,N=%d,%s]", BN, D, D, kPrec);
  float *t6 = s1, *m_e = s2, *dy3 = s3, *dz4 = s4, *dA = s5;
  if (live) {
    mlp_recompute(q, y3, dyo, P, FC1E_W, LN6_W, t6, io[DG_BLK_E_H], io[DG_BLK_E_MASK], m_e, dy3, io[DG_BLK_E_H2], R, D, H, eps, ws, wsb);
    DG_STEP(q, dg_add_ln_bwd(dy3, z4, nullptr, P[LN4_W], dz4, scratch, scratch + D, R, D, eps, 0, q.s), "add_ln_bwd");
    DG_STEP(q, dg_rows_gemm(dz4, P[OE_W], 0, nullptr, 0, nullptr, nullptr, dA, R, D, D, kP, 0, q.s), "rows_gemm[R=%lld,K=%d,N=%d,%s]", R, D, D, kPrec);
  }
  zero(q, ns(nDV), nb);
  DG_STEP(q, dg_softmax_agg_bwd(ns(nDG), a4, vv, dA, ns(nDV), live ? 1 : 0, B, N, D, q.s), "softmax_agg_bwd");      // dA: out_e path + softmax path
  zero(q, ns(nDK), nb);
  float* dE = s6;
  DG_STEP(q, dg_modulate_bwd(dA, qq, kk, e, c, ns(nDQ), ns(nDK), dE, B, N, D, q.s), "modulate_bwd");
  const float* dx1 = ns(nDZ3);
  {
    const int dt[3] = {nDQ, nDK, nDV}, out[3] = {nP0, nP1, nP0};
    for (int i = 0; i < 3; ++i) {
      DG_STEP(q, dg_rows_gemm(ns(dt[i]), P[w3[i]], 0, nullptr, 0, nullptr, dx1, ns(out[i]), BN, D, D, kP, 0, q.s),
              "rows_gemm[R=%lld,K=%d,N=%d,%s+resid]", BN, D, D, kPrec);
      dx1 = ns(out[i]);
    }
  }
  // ---- 2. reverse of the first-order backward program
  float* c_x = at(io, DG_BLK_C_X);
  DG_STEP(q, dg_add_ln_bwd_bwd(ux, nullptr, nullptr, dx1, x, nullptr, P[LN1_W], ns(nCDX1), c_x, grads[LN1_W], BN, D, eps, q.s), "add_ln_bwd_bwd");
  {
    const int dt[3] = {nDQ, nDK, nDV}, cd[3] = {nCDQ, nCDK, nCDV};
    for (int i = 0; i < 3; ++i) {                       // dx1 = dz3 + dq Wq + dk Wk + dv Wv
      DG_STEP(q, dg_rows_gemm(ns(nCDX1), P[w3[i]], 1, nullptr, 0, nullptr, nullptr, ns(cd[i]), BN, D, D, kP, 0, q.s),
              "rows_gemm[R=%lld,K=%d,N=%d,%s]", BN, D, D, kPrec);
      DG_STEP(q, dg_gemm_tn(ns(dt[i]), ns(nCDX1), grads[w3[i]], nullptr, BN, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", BN, D, D, kPrec);
    }
  }
  float* c_dE = s7;                                      // dy = dz4 + dE We
  DG_STEP(q, dg_rows_gemm(uy, P[E_W], 1, nullptr, 0, nullptr, nullptr, c_dE, R, D, D, kP, 0, q.s), "rows_gemm[R=%lld,K=%d,N=%d,%s]", R, D, D, kPrec);
  DG_STEP(q, dg_gemm_tn(dE, uy, grads[E_W], nullptr, R, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", R, D, D, kPrec);
  float *c_dA = s6, *c_E = s8;                           // (dE is dead: its slot takes c[dA])
  zero(q, ns(nCK), nb);
  DG_STEP(q, dg_modulate_bwd_bwd(ns(nCDQ), ns(nCDK), c_dE, dA, qq, kk, e, c, c_dA, ns(nCQ), ns(nCK), c_E, B, N, D, q.s), "modulate_bwd_bwd");
  float* c_A = s5;                                       // (dA is dead)
  zero(q, ns(nCV), nb);
  DG_STEP(q, dg_softmax_agg_bwd_bwd(c_dA, ns(nCDV), ns(nDG), a4, vv, ns(nCDG), c_A, ns(nCV), B, N, D, q.s), "softmax_agg_bwd_bwd");
  float *c_z4 = nullptr, *c_me = nullptr;
  if (live) {
    float* c_dz4 = s7;                                   // da = dz4 Woe;  dy = dz4 + ...   (c[dE] is dead)
    DG_STEP(q, dg_rows_gemm(c_dA, P[OE_W], 1, nullptr, 0, nullptr, uy, c_dz4, R, D, D, kP, 0, q.s), "rows_gemm[R=%lld,K=%d,N=%d,%s+resid]", R, D, D, kPrec);
    DG_STEP(q, dg_gemm_tn(dz4, c_dA, grads[OE_W], nullptr, R, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", R, D, D, kPrec);
    float* c_dy3 = s4;                                   // (dz4 and c[dA] are dead)
    c_z4 = s6;
    DG_STEP(q, dg_add_ln_bwd_bwd(c_dz4, nullptr, nullptr, dy3, z4, nullptr, P[LN4_W], c_dy3, c_z4, grads[LN4_W], R, D, eps, q.s), "add_ln_bwd_bwd");
    float* c_t6 = s3;                                    // (dy3 is dead)
    mlp_second(q, c_dy3, t6, io[DG_BLK_E_MASK], io[DG_BLK_E_H2], P, FC1E_W, grads, c_t6, io[DG_BLK_E_H3], wt, R, D, H, ws, wsb);
    c_me = s1;                                           // (t6 is dead)
    DG_STEP(q, dg_add_ln_bwd_bwd(c_t6, nullptr, nullptr, dyo, m_e, nullptr, P[LN6_W], at(io, DG_BLK_C_DYO), c_me, grads[LN6_W], R, D, eps, q.s),
            "add_ln_bwd_bwd");
  }
  // dg = dz3 Won;  dx1 = dz3 + ...
  DG_STEP(q, dg_rows_gemm(ns(nCDG), P[ON_W], 1, nullptr, 0, nullptr, ns(nCDX1), ns(nCDZ3), BN, D, D, kP, 0, q.s),
          "rows_gemm[R=%lld,K=%d,N=%d,%s+resid]", BN, D, D, kPrec);
  DG_STEP(q, dg_gemm_tn(ns(nDZ3), ns(nCDG), grads[ON_W], nullptr, BN, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", BN, D, D, kPrec);
  DG_STEP(q, dg_add_ln_bwd_bwd(ns(nCDZ3), nullptr, nullptr, ns(nDX3), x1, on, P[LN3_W], ns(nCDX3), ns(nCZ3), grads[LN3_W], BN, D, eps, q.s),
          "add_ln_bwd_bwd");
  mlp_second(q, ns(nCDX3), ns(nT5), io[DG_BLK_N_MASK], io[DG_BLK_N_H2], P, FC1_W, grads, ns(nCT5), io[DG_BLK_N_H3], wt, BN, D, H, ws, wsb);
  DG_STEP(q, dg_add_ln_bwd_bwd(ns(nCT5), nullptr, nullptr, dxo, ns(nM), nullptr, P[LN5_W], at(io, DG_BLK_C_DXO), ns(nCMN), grads[LN5_W], BN, D, eps,
                               q.s),
          "add_ln_bwd_bwd");
  // ---- 3. first-order backward of the forward program from the injected cotangents
  if (live) {
    float* c_y3 = s2;                                    // (m_e is dead)
    mlp_first(q, c_me, io[DG_BLK_E_H], io[DG_BLK_E_MASK], y3, P, FC1E_W, grads, c_y3, io[DG_BLK_E_H3], R, D, H, ws, wsb);
    DG_STEP(q, dg_add_ln_bwd(c_y3, z4, nullptr, P[LN4_W], c_z4, grads[LN4_W], grads[LN4_B], R, D, eps, 1, q.s), "add_ln_bwd");   // c[z4] += LN4^T c[y3]
    float* c_A2 = s4;                                    // y1 = A Woe^T + boe   (c[dy3] is dead)
    DG_STEP(q, dg_rows_gemm(c_z4, P[OE_W], 0, nullptr, 0, nullptr, c_A, c_A2, R, D, D, kP, 0, q.s), "rows_gemm[R=%lld,K=%d,N=%d,%s+resid]", R, D, D, kPrec);
    DG_STEP(q, dg_gemm_tn(c_z4, a4, grads[OE_W], grads[OE_B], R, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", R, D, D, kPrec);
    c_A = c_A2;
  }
  mlp_first(q, ns(nCMN), io[DG_BLK_N_H], io[DG_BLK_N_MASK], x3, P, FC1_W, grads, ns(nCX3), io[DG_BLK_N_H3], BN, D, H, ws, wsb);
  DG_STEP(q, dg_add_ln_bwd(ns(nCX3), x1, on, P[LN3_W], ns(nCZ3), grads[LN3_W], grads[LN3_B], BN, D, eps, 1, q.s), "add_ln_bwd");
  DG_STEP(q, dg_rows_gemm(ns(nCZ3), P[ON_W], 0, nullptr, 0, nullptr, nullptr, ns(nCG), BN, D, D, kP, 0, q.s), "rows_gemm[R=%lld,K=%d,N=%d,%s]", BN, D, D, kPrec);
  DG_STEP(q, dg_gemm_tn(ns(nCZ3), g, grads[ON_W], grads[ON_B], BN, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", BN, D, D, kPrec);
  // attention backward from (c[g], c[A]): c[E] += in the store; its dq / dk / dv (the ring kernel STORES dk and dv) are added to
  // c[q] / c[k] / c[v] in one small launch
  zero(q, ns(nDQ2), 3 * nb);
  DG_STEP(q, dg_attn_scores_bwd(ns(nCG), c_A, qq, kk, vv, e, c, sm, si, g, c_E, ns(nDQ2), ns(nDK2), ns(nDV2), B, N, D, 8, q.s), "attn_scores_bwd[fused]");
  if (!q.rc && dg::trace_on()) {
    dg::trace_call("add3", (const void*)ns(nCQ), (const void*)ns(nDQ2), (const void*)ns(nCK), (const void*)ns(nDK2), (const void*)ns(nCV),
                   (const void*)ns(nDV2), BN * D);
  } else if (!q.rc) {
    const long long n4 = BN * D / 4;
    const long long blocks = (n4 + 255) / 256, cap = (long long)dg::sm_count() * 8;
    add3_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, q.s>>>(reinterpret_cast<float4*>(ns(nCQ)), reinterpret_cast<const float4*>(ns(nDQ2)),
                                                                     reinterpret_cast<float4*>(ns(nCK)), reinterpret_cast<const float4*>(ns(nDK2)),
                                                                     reinterpret_cast<float4*>(ns(nCV)), reinterpret_cast<const float4*>(ns(nDV2)), n4);
    q.rc = dg::check_launch("dg_block_bwd_bwd(add3)");
  }
  // E = y We^T + be;  z4 = y + y1
  DG_STEP(q, dg_rows_gemm(c_E, P[E_W], 0, nullptr, 0, nullptr, live ? c_z4 : nullptr, io[DG_BLK_C_Y], R, D, D, kP, 0, q.s),
          "rows_gemm[R=%lld,K=%d,N=%d,%s%s]", R, D, D, kPrec, live ? "+resid" : "");
  DG_STEP(q, dg_gemm_tn(c_E, y, grads[E_W], grads[E_B], R, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", R, D, D, kPrec);
  const float* c_x1 = ns(nCZ3);
  {
    const int ct[3] = {nCQ, nCK, nCV}, out[3] = {nP0, nP1, nP0};
    for (int i = 0; i < 3; ++i) {
      DG_STEP(q, dg_gemm_tn(ns(ct[i]), x1, grads[w3[i]], grads[w3[i] + 1], BN, D, D, kP, 0, q.s), "gemm_tn[R=%lld,M=%d,N=%d,%s]", BN, D, D, kPrec);
      DG_STEP(q, dg_rows_gemm(ns(ct[i]), P[w3[i]], 0, nullptr, 0, nullptr, c_x1, ns(out[i]), BN, D, D, kP, 0, q.s),
              "rows_gemm[R=%lld,K=%d,N=%d,%s+resid]", BN, D, D, kPrec);
      c_x1 = ns(out[i]);
    }
  }
  DG_STEP(q, dg_add_ln_bwd(c_x1, x, nullptr, P[LN1_W], c_x, grads[LN1_W], grads[LN1_B], BN, D, eps, 1, q.s), "add_ln_bwd");
  return q.done();
}

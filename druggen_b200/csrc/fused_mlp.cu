// Fused residual-MLP chain of the encoder block (layers.py:41-54 + the residual and LayerNorm of
// layers.py:191/192) as ONE tcgen05 kernel skeleton with four epilogue personalities:
//
//   FWD    out = LayerNorm( x + fc2( relu( fc1(x) + b1 ) ) + b2 ) * gamma + beta
//   BWD_A  recompute h = relu(fc1(x)+b1), z = x + fc2(h) + b2, then the LayerNorm BACKWARD of `dout`
//          through z:  out = dz [R,128] fp32,  spill = h [R,H] bf16,  dgamma/dbeta += column sums
//   BWD_B  dh = (dz . W2) * (h > 0)   (spilled as bf16 for the weight-gradient pass),
//          out = dx = dz + dh . W1
//   ATTN   the edge half of the attention (layers.py:116,123-127 + the residual and LayerNorm of :188/:190):
//          E = y . We^T + be;  A = c q_i k_j (E^2 + E);  out = LN4( y + A . Woe^T + boe ) -- one hidden chunk (H = 128),
//          "fc1" = We with the modulation as its epilogue, "fc2" = Woe.  Optional side outputs: A as bf16 (operand of the
//          softmax-aggregate and of dWoe), E fp32 and the pre-LayerNorm sum fp32 (for the backward).
//
// All four are "GEMM1 per 128-wide hidden chunk -> per-row epilogue -> bf16 operand block in smem -> GEMM2
// accumulating over chunks -> per-row final epilogue": the 128 x H intermediate never round-trips HBM as
// fp32 (it leaves once as bf16 -- TMA-stored straight from the operand block -- only where something downstream needs it).
// The final epilogue's rows come in (residual / upstream gradient) and go out as 128-byte-swizzled TMA boxes that each
// epilogue warp loads, rewrites in place and stores for its own 32 rows.
//
// Persistent CTA per SM, 704 threads, warp-specialised:
//   warps 0-15  epilogue : thread = a quarter of a row (TMEM lane quarter w&3, 32-column part w>>2); one TMA box
//                          [32 rows][32 channels] of the I/O tile per warp, bf16 side outputs stored per warp PAIR
//   warps 16-19 loader   : fp32 LDG.E.256 -> bf16 -> swizzled operand blocks (x tile and operand block single-buffered; the
//                          64 KB I/O tile of the final epilogue is dedicated in every mode, see kDedIO)
//   warp  20    MMA      : one thread issues tcgen05.mma; GEMM1 of tile t+1 issued right after the last GEMM2 of tile t
//   warp  21    W loader : one thread streams pre-packed bf16 weight stages (32 KB) with
//                          cp.async.bulk into a 2-stage ring (weights live in L2: 192 KB per net; ATTN: both stages resident)
// TMEM: columns [0,384) three GEMM1 chunk accumulators, [384,512) the GEMM2 accumulator.
#include <cuda.h>
#include "tc_common.cuh"
#include "../../include/druggen_b200.h"

#ifndef DG_CHAIN_DEDICATED_IO
#define DG_CHAIN_DEDICATED_IO 1
#endif

namespace dg {
namespace tc {

// Epilogue geometry: 16 warps = TMEM lane quarter (w & 3: rows 32q..32q+31 of the tile) x column part (w >> 2: 32 of the 128
// columns).  A thread = (row, 32 columns): half the dependent work per phase of the former 8-warp / 64-column layout, whose
// serial per-tile epilogue (not the tensor pipe, not HBM) bounded all four kernels (profiles/r02a_chain_phases.jsonl).
constexpr int kEpiWarps = 16;
constexpr int kParts = kEpiWarps / 4;            // column parts of a row
constexpr int kCW = 128 / kParts;                // columns per epilogue thread = one TMA box [32 rows][32 channels] per warp
constexpr int kEpiThreads = kEpiWarps * 32;      // 512
constexpr int kLdWarp0 = kEpiWarps;              // loader warps 16..19
constexpr int kMmaWarp = kEpiWarps + 4;          // 20
constexpr int kWWarp = kEpiWarps + 5;            // 21
constexpr int kMlpThreads = (kEpiWarps + 6) * 32;   // 704
static_assert(kCW == 32, "the epilogue below is written for 32 columns per thread");
constexpr int kWStage = 2 * kBlkBytes;   // one packed weight stage: [2 kb][128 rows][128 B] = 32 KB
constexpr int kStgPitch = 8;             // epilogue transpose: 32 rows x 8 words per warp, XOR-swizzled (no padding)
// word offset of the 4-float chunk c2 (0..1) of row r in a staging tile: conflict-free both for the coalesced side (8 lanes = 4
// rows x 2 chunks) and for the thread-per-row side (8 lanes = 8 consecutive rows, same chunk)
__device__ __forceinline__ int stg_off(int r, int c2) { return r * kStgPitch + ((c2 ^ ((r >> 2) & 1)) << 2); }

enum { kFwd = 0, kBwdA = 1, kBwdB = 2, kAttn = 3 };

struct MlpArgs {
  const float* x;          // FWD/BWD_A: block input [R,128];  BWD_B: dz [R,128]
  const uint8_t* wpack;    // packed bf16 weight stages
  const float* b1;         // [H]      (FWD/BWD_A)
  const float* b2;         // [128]    (FWD/BWD_A)
  const float* gamma;      // [128]    (FWD/BWD_A)
  const float* beta;       // [128]    (FWD)
  const float* dout;       // BWD_A: upstream gradient [R,128]
  float* out;              // FWD: y;  BWD_A: dz;  BWD_B: dx     [R,128]
  uint16_t* spill;         // BWD_A: h out [R,H] bf16;  BWD_B: dh out [R,H] bf16
  const uint16_t* gate;    // BWD_B: h in [R,H] bf16 (sign mask source when mask_in is NULL)
  unsigned long long* mask_out;       // BWD_A: optional ReLU sign mask [R][H/64] (bit i of word (c*2+hf): h[c*128+hf*64+i] > 0)
  const unsigned long long* mask_in;  // BWD_B: optional, replaces the 2 H bytes per row of `gate` by H/8
  float* dgamma;           // BWD_A: += [128]
  float* dbeta;            // BWD_A: += [128]
  long long R;
  int HC;
  float eps;
  // ---- ATTN only
  const float* q;          // [B*N,128]
  const float* k;          // [B*N,128]
  float* e_out;            // optional E [R,128] fp32
  float* z_out;            // optional y + out_e(A) [R,128] fp32 (input of LN4)
  int natoms;              // N: edge row r = ((b N + i) N + j)
  float cscale;            // 1 / sqrt(d_k)
  int prefetch;            // DG_OPT_L2_PREFETCH
  long long* prof;         // debug: per-CTA phase cycle counters [grid][4 roles][16] (dg_debug_chain_profile), or NULL
  // tensor maps of the [R,128] fp32 row tensors that cross the final epilogue through TMA: box = [128 rows][32 channels],
  // 128-byte swizzle (tm_x: residual input = x;  tm_out: out;  tm_z: z_out of ATTN)
  alignas(64) CUtensorMap tm_x;
  alignas(64) CUtensorMap tm_out;
  alignas(64) CUtensorMap tm_z;
  alignas(64) CUtensorMap tm_e;      // ATTN: e_out
  alignas(64) CUtensorMap tm_dout;   // BWD_A: the upstream gradient rows
  // the bf16 side output [R,H] (h / dh / the scores): box = [128 rows][64 channels] = one operand block, 128-byte swizzle --
  // it leaves straight out of the GEMM2 operand buffer the epilogue has just filled
  alignas(64) CUtensorMap tm_spill;
};

// phase timing (debug): lane 0 of four role-leader warps accumulates clock64() deltas per phase into shared memory
#define DG_PROF(id)                                            \
  if (prof_on) {                                              \
    const long long t_now = clock64();                        \
    sProf[prof_slot * 16 + (id)] += t_now - t_prev;           \
    t_prev = t_now;                                           \
  }

// ---- weight pre-pack: fp32 nn.Linear weights -> bf16 swizzled operand stages in a workspace ------
// forward orientation (FWD, BWD_A):
//   stage c      : fc1 rows [c*128, +128) of W1[H,128]            B operand: N = hidden unit, K = in
//   stage HC + c : fc2 columns [c*128, +128) of W2[128,H]         B operand: N = out,         K = hidden slice
// transposed orientation (BWD_B):
//   stage c      : W2^T slice: element (n, o) = W2[o, c*128+n]    B operand: N = hidden unit, K = out
//   stage HC + c : W1^T slice: element (k, n) = W1[c*128+n, k]    B operand: N = in,          K = hidden slice
__global__ void mlp_pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2, uint8_t* __restrict__ ws,
                                        int H, int transposed) {
  const int HC = H / 128;
  const int total = 2 * HC * 128 * 16;          // (stage, row, kb*8 + j)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int c16 = idx & 15, row = (idx >> 4) & 127, stage = idx >> 11;
    int kb = c16 >> 3, j = c16 & 7;
    const int k0 = kb * 64 + j * 8;              // first of 8 consecutive K indices
    float4 lo, hi;
    if (!transposed) {
      const float* src = stage < HC ? w1 + (long long)(stage * 128 + row) * 128 + k0
                                    : w2 + (long long)row * H + (stage - HC) * 128 + k0;
      lo = ld4(src); hi = ld4(src + 4);
    } else {
      float t[8];
      if (stage < HC) {                          // (n=row, o=k0+i) = W2[o, stage*128 + n]
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = w2[(long long)(k0 + i) * H + stage * 128 + row];
      } else {                                   // (k=row, n=k0+i) = W1[(stage-HC)*128 + n, k]
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = w1[(long long)((stage - HC) * 128 + k0 + i) * 128 + row];
      }
      lo = make_float4(t[0], t[1], t[2], t[3]); hi = make_float4(t[4], t[5], t[6], t[7]);
    }
    st_block_chunk(ws + (long long)stage * kWStage + kb * kBlkBytes, row, j, lo, hi);
  }
}

struct MlpSmem {
  // tile buffers, 32 KB units: x (single-buffered) | operand block | weight stages 2 x 32 KB | the dedicated 64 KB I/O tile
  // (DG_CHAIN_DEDICATED_IO=0: two operand buffers in the H = 384 modes, which then also stage their I/O tile)
  static constexpr int tiles = 0;
  static constexpr int stage = 6 * kWStage;                          // 16 warps x 32 x kStgPitch words
  static constexpr int stats = stage + kEpiWarps * 32 * kStgPitch * 4;   // 2 x [4 parts][128 rows] float2 (see the epilogue)
  static constexpr int vec = stats + 2 * kParts * 128 * 8;           // b1[384] b2[128] gamma[128] beta[128]
  static constexpr int bars = vec + (384 + 3 * 128) * 4;
  static constexpr int prof = bars + 512;                            // 4 roles x 16 counters (debug)
  static constexpr int total = prof + 4 * 16 * 8;
};

// warp-cooperative transposes through a [32][kStgPitch] staging tile, 8 words (columns) at a time:
// global rows <-> "thread = row" registers.  Global accesses are 32 B contiguous per row (pitch in words); the coalesced side
// is lane = (row it * 16 + (lane >> 1), 16-byte half lane & 1).
__device__ __forceinline__ void gather_issue(const float* __restrict__ src, long long row_base, long long R, int pitch, int col,
                                             int nrounds, int lane, float4* xq) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (g < nrounds) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = it * 16 + (lane >> 1);
        xq[g * 2 + it] = (row_base + r < R) ? ld4(src + (row_base + r) * pitch + col + g * 8 + (lane & 1) * 4)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
}
__device__ __forceinline__ void gather_finish(const float4* xq, int nrounds, float* stg, int lane, float* dst) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (g < nrounds) {
#pragma unroll
      for (int it = 0; it < 2; ++it) st4(stg + stg_off(it * 16 + (lane >> 1), lane & 1), xq[g * 2 + it]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = ld4(stg + stg_off(lane, i));
        dst[g * 8 + 4 * i] = v.x; dst[g * 8 + 4 * i + 1] = v.y; dst[g * 8 + 4 * i + 2] = v.z; dst[g * 8 + 4 * i + 3] = v.w;
      }
      __syncwarp();
    }
}

// Column sums over the warp's 32 rows without shared memory (BWD_A's dgamma / dbeta): three recursive-halving exchanges over lane
// bits 4, 3, 2 leave lane L with the sum over 8 rows of column 4 b4 + 2 b3 + b2 of an 8-column group (7 shuffles for 8 columns);
// the remaining two lane bits are folded once, when the kernel flushes.  Replaces a staged transpose per group and tile (two
// warp barriers and four shared-memory round trips each: ~3 k of the tile's ~27 k cycles) and 12 accumulator registers.
__device__ __forceinline__ float colsum8(const float (&p)[8], int lane) {
  const bool u4 = lane & 16, u3 = lane & 8, u2 = lane & 4;
  float q[4], r[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = (u4 ? p[4 + i] : p[i]) + __shfl_xor_sync(0xffffffffu, u4 ? p[i] : p[4 + i], 16);
#pragma unroll
  for (int i = 0; i < 2; ++i) r[i] = (u3 ? q[2 + i] : q[i]) + __shfl_xor_sync(0xffffffffu, u3 ? q[i] : q[2 + i], 8);
  return (u2 ? r[1] : r[0]) + __shfl_xor_sync(0xffffffffu, u2 ? r[0] : r[1], 4);
}

// 22 warps = 6 on two of the SM's four register-file partitions (16 384 registers each): 80 registers per thread is the ceiling
// (6 x 32 x 88 does not fit -- "too many resources requested"), which ptxas picks under this launch bound.
template <int kMode>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_chain_tc_kernel(const __grid_constant__ MlpArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: a pointer rebuilt from an integer loses its address
  // space and every access through it compiles to generic LD/ST (LSU long-scoreboard path) instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // ATTN has ONE hidden chunk and two weight stages in all: both weight stages stay resident, the operand block is
  // single-buffered, and the 64 KB I/O tile is DEDICATED, so the residual rows of a tile are requested a whole chunk phase
  // before they are needed (in the other modes the I/O tile aliases the two operand buffers and can only be requested after
  // the tile's last GEMM2: ~2.5 k cycles of exposed TMA latency per tile).  The x tile is single-buffered in every mode: the
  // loader has a whole tile period between GEMM1 of tile t and GEMM1 of tile t+1 (issued after tile t's last GEMM2).
  // Round 2 (DG_CHAIN_DEDICATED_IO, default on): the H = 384 modes use five of the six 32 KB tile units -- x, two operand buffers,
  // two weight stages.  Single-buffering the operand block as ATTN does (its reuse wait hides behind the next chunk's
  // accumulator read and math: kLateHb) frees the unit that, with the sixth, makes the 64 KB I/O tile DEDICATED in every mode: the
  // residual / dout box of a tile is requested right after the tile's FIRST chunk instead of after its last GEMM2.
  constexpr bool kResW = kMode == kAttn;                         // both weight stages stay resident (one hidden chunk)
  constexpr bool kDedIO = kMode == kAttn || DG_CHAIN_DEDICATED_IO;
  constexpr bool kLateHb = kDedIO && kMode != kAttn;
  uint8_t* sX = smem;
  uint8_t* sH = smem + kWStage;
  uint8_t* sW = smem + (kDedIO ? 2 : 3) * kWStage;
  uint8_t* sIO = kDedIO ? smem + 4 * kWStage : sH;
  float* sStage = reinterpret_cast<float*>(smem + MlpSmem::stage);
  float2* sStats = reinterpret_cast<float2*>(smem + MlpSmem::stats);
  float* sB1 = reinterpret_cast<float*>(smem + MlpSmem::vec);
  float* sB2 = sB1 + 384;
  float* sG = sB2 + 128;
  float* sBe = sG + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MlpSmem::bars);
  uint64_t *x_full = bars, *x_empty = bars + 1, *w_full = bars + 2, *w_empty = bars + 4, *hacc_full = bars + 6,
           *hacc_empty = bars + 9, *hb_full = bars + 12, *hb_empty = bars + 14, *z_full = bars + 16, *z_empty = bars + 17,
           *sp_done = bars + 18, *io_full = bars + 19 /* [kEpiWarps]: one per epilogue warp */;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19 + kEpiWarps);
  long long* sProf = reinterpret_cast<long long*>(smem + MlpSmem::prof);

  const float* __restrict__ x = A.x;
  const long long R = A.R;
  const int HC = A.HC, H = HC * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long num_tiles = (R + 127) / 128;
  const long long my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const bool spill = (kMode == kBwdA || kMode == kBwdB || kMode == kAttn) && A.spill != nullptr;
  const bool use_mask = kMode == kBwdB && A.mask_in != nullptr;
  const bool want_affine = kMode == kBwdA && A.dgamma != nullptr;      // BWD_A: the LayerNorm's dgamma / dbeta are wanted

  if (tid == 0) {
    mbar_init(x_full, 128); mbar_init(x_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
      mbar_init(&hb_full[i], kEpiWarps); mbar_init(&hb_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) { mbar_init(&hacc_full[i], 1); mbar_init(&hacc_empty[i], kEpiWarps); }
    mbar_init(z_full, 1); mbar_init(z_empty, kEpiWarps);     // epilogue barriers: one elected arrival per warp
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&io_full[i], 1);
    mbar_init(sp_done, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  if (tid < 64) sProf[tid] = 0;
  // (only epilogue warps 0 and 4 -- column parts 0 and 1 of lane quarter 0 --, the first loader warp and the MMA warp record)
  const int prof_slot = warp == 0 ? 0 : warp == 4 ? 1 : warp == kLdWarp0 ? 2 : 3;
  const bool prof_on = A.prof != nullptr && lane == 0 && (warp == 0 || warp == 4 || warp == kLdWarp0 || warp == kMmaWarp);
  long long t_prev = clock64();
  if (kMode != kBwdB) {
    for (int i = tid; i < H; i += kMlpThreads) sB1[i] = A.b1[i];
    for (int i = tid; i < 128; i += kMlpThreads) {
      sB2[i] = A.b2[i];
      sG[i] = A.gamma[i];
      sBe[i] = (kMode == kFwd || kMode == kAttn) ? A.beta[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kLdWarp0 && warp < kLdWarp0 + 4) {
    // ------------------------------------------------------------------ input-tile loader
    const int lt = tid - kEpiThreads;
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long row0 = (blockIdx.x + ti * gridDim.x) * 128;
      if ((A.prefetch & 1) && (lt == 0 || lt == 32)) {
        // TMA-engine L2 prefetch of whole (contiguous) row tiles ahead of the register-staged loads:
        // thread 0: the input tiles ti+1, ti+2;  thread 32: what the epilogue gathers for tile ti+1 (dout / bf16 gate)
        for (long long tp = ti + 1; tp <= ti + 2; ++tp) {
          if (tp >= my_tiles || (tp == ti + 1 && ti != 0 && lt == 0) || (tp == ti + 2 && lt == 32)) continue;
          const long long prow0 = (blockIdx.x + tp * gridDim.x) * 128;
          const long long prows = R - prow0 < 128 ? R - prow0 : 128;
          if (lt == 0) bulk_prefetch_l2(x + prow0 * 128, prows * 512);
          else if (kMode == kBwdA) bulk_prefetch_l2(A.dout + prow0 * 128, prows * 512);
          else if (kMode == kBwdB && !use_mask) bulk_prefetch_l2(A.gate + prow0 * H, prows * H * 2);
        }
      }
      DG_PROF(0)
      mbar_wait(x_empty, (ti & 1) ^ 1);
      DG_PROF(1)
      // this thread: rows (lt >> 3) + 16 it, 32-byte chunk lt & 7 of each 64-channel half -- one base pointer, immediate offsets
      const float* xt = x + row0 * 128 + (lt >> 3) * 128 + (lt & 7) * 8;
      const bool full_tile = row0 + 128 <= R;          // (every tile but the last: no row predicates in the load burst)
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        uint8_t* blk = sX + kb * kBlkBytes;
        float4 v[16];
        if (full_tile && !(A.prefetch & 2)) {
#pragma unroll
          for (int it = 0; it < 8; ++it) ld8(xt + it * 2048 + kb * 64, v[2 * it], v[2 * it + 1]);
        } else {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 16 + (lt >> 3);
            if (row0 + r < R) {
              if (A.prefetch & 2) ld8_keep(xt + it * 2048 + kb * 64, v[2 * it], v[2 * it + 1]);   // x is re-read as the residual
              else ld8(xt + it * 2048 + kb * 64, v[2 * it], v[2 * it + 1]);
            } else {
              v[2 * it] = v[2 * it + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          int item = it * 128 + lt;
          st_block_chunk(blk, item >> 3, item & 7, v[2 * it], v[2 * it + 1]);
        }
      }
      fence_async_smem();
      mbar_arrive(x_full);
      DG_PROF(2)
    }
  } else if (warp == kWWarp) {
    // ------------------------------------------------------------------ weight streamer (one thread)
    if (lane == 0 && my_tiles > 0) {
      uint32_t wcount = 0;
      auto push = [&](int stage) {
        const int ws = wcount & 1;
        mbar_wait(&w_empty[ws], ((wcount >> 1) & 1) ^ 1);
        mbar_expect_tx(&w_full[ws], kWStage);
        bulk_g2s(sW + ws * kWStage, A.wpack + (long long)stage * kWStage, kWStage, &w_full[ws]);
        ++wcount;
      };
      for (int c = 0; c < HC; ++c) push(c);                         // GEMM1 of the first tile
      if (kResW) push(1);                                           // ATTN: stage 1 (out_e) into slot 1 -- both stay resident
      // consumption order per tile: GEMM2(0..HC-1), then GEMM1 of the next tile (0..HC-1)
      for (long long ti = 0; !kResW && ti < my_tiles; ++ti) {
        for (int c = 0; c < HC; ++c) push(HC + c);
        if (ti + 1 < my_tiles)
          for (int c = 0; c < HC; ++c) push(c);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = make_idesc(128, 128, 0, 0);
      uint32_t wcount = 0, hcount = 0;
      auto mma_chunk = [&](uint32_t a_base, uint32_t d_col, bool first_clears) {
        const int ws = kResW ? (d_col == 384 ? 1 : 0) : (wcount & 1);       // ATTN: resident stages, slot = which GEMM
        DG_PROF(0)
        mbar_wait(&w_full[ws], kResW ? 0 : ((wcount >> 1) & 1));
        DG_PROF(1)
        tc_fence_after();
        const uint32_t b_base = smem_u32(sW + ws * kWStage);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * kBlkBytes + (kk & 3) * 32;
          umma_bf16(tmem_base + d_col, make_sdesc(a_base + off, 16, 1024), make_sdesc(b_base + off, 16, 1024), idesc,
                    (first_clears && kk == 0) ? 0u : 1u);
        }
        if (!kResW) umma_commit(&w_empty[ws]);
        ++wcount;
      };
      auto gemm1 = [&](long long ti, int c) {
        DG_PROF(0)
        if (c == 0) { mbar_wait(x_full, ti & 1); tc_fence_after(); }
        DG_PROF(2)
        mbar_wait(&hacc_empty[c], (ti & 1) ^ 1);
        DG_PROF(3)
        tc_fence_after();
        mma_chunk(smem_u32(sX), c * 128, true);
        umma_commit(&hacc_full[c]);
        if (c == HC - 1) umma_commit(x_empty);
      };
      for (int c = 0; c < HC; ++c) gemm1(0, c);
      for (long long ti = 0; ti < my_tiles; ++ti) {
        for (int c = 0; c < HC; ++c) {
          const int hs = kDedIO ? 0 : (hcount & 1);
          DG_PROF(0)
          mbar_wait(&hb_full[hs], kDedIO ? (hcount & 1) : ((hcount >> 1) & 1));
          DG_PROF(4)
          if (c == 0) mbar_wait(z_empty, (ti & 1) ^ 1);
          DG_PROF(5)
          tc_fence_after();
          mma_chunk(smem_u32(sH + hs * kWStage), 384, c == 0);
          umma_commit(&hb_empty[hs]);
          ++hcount;
          if (c == HC - 1) umma_commit(z_full);      // before the next tile's GEMM1 is even issued: the final epilogue does not wait for it
        }
        // GEMM1 of the next tile runs while the epilogue warps are in this tile's final epilogue
        if (ti + 1 < my_tiles)
          for (int c = 0; c < HC; ++c) gemm1(ti + 1, c);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-15)
    // warp w: TMEM lane quarter q = w & 3 (rows 32q..32q+31 of the tile), column part cp = w >> 2 (columns 32cp..32cp+31)
    const int q = warp & 3, cp = warp >> 2;
    float* stg = sStage + warp * 32 * kStgPitch;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    // the bf16 side output is stored per operand block = 64 columns = TWO warps' rows: the pair (cp even, cp odd) of a lane
    // quarter meets at its own named barrier and the even one issues the TMA store
    const bool issuer = (cp & 1) == 0;
    const int pair_bar_id = 2 + q + 4 * (cp >> 1);
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar_id) : "memory"); };
    uint32_t hcount = 0;
    // BWD_A column sums, kept per lane over all tiles (colsum8): lane = (column 4 b4 + 2 b3 + b2 of each 8-column group, rows with
    // this lane's two low bits)
    float acc_g[4] = {0.f, 0.f, 0.f, 0.f}, acc_b[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long row0 = (blockIdx.x + ti * gridDim.x) * 128;
      const long long wrow0 = row0 + q * 32;                          // first global row of this warp
      uint8_t* iobox = sIO + cp * kBlkBytes + q * 32 * 128;           // this warp's box [32 rows][32 channels] of the I/O tile
      uint8_t* iorow = sIO + cp * kBlkBytes + row * 128;              // this thread's row of it
      auto io_chunk = [&](int j) -> float* {                          // 16-byte chunk j (4 channels) of this thread's 32 columns
        return reinterpret_cast<float*>(iorow + ((j ^ (row & 7)) << 4));
      };
      if (kMode == kBwdA && wrow0 + lane < R)                         // the LayerNorm-backward's dout rows come from L2
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.dout + (wrow0 + lane) * 128 + cp * kCW));
      for (int c = 0; c < HC; ++c) {
        float4 gq[4];
        uint32_t mk = 0u;                              // BWD_B: this thread's 32 ReLU sign bits of chunk c (one 4-byte load)
        if (kMode == kBwdB) {
          if (use_mask) {
            if (wrow0 + lane < R) mk = __ldg(reinterpret_cast<const uint32_t*>(A.mask_in) + (wrow0 + lane) * (4 * HC) + c * 4 + cp);
          } else {   // bf16 [R,H] viewed as 32-bit words: pitch H/2, this thread's 32 columns = 16 words
            gather_issue(reinterpret_cast<const float*>(A.gate), wrow0, R, H >> 1, (c * 128 + cp * kCW) >> 1, 2, lane, gq);
          }
        }
        // ATTN: q / k rows of the 2 staged rows this lane serves (row r = ((b N + i) N + j) -> q row b N + i, k row b N + j),
        // and the first 8-channel group of both, requested before the accumulator is waited for
        unsigned qoff[2], koff[2];                     // element offsets (B N 128 < 2^31)
        float4 qc[2], kc[2];
        if (kMode == kAttn) {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            long long r = wrow0 + it * 16 + (lane >> 1);
            if (r >= R) r = R - 1;
            const unsigned n = (unsigned)A.natoms, bi = (unsigned)r / n, jj = (unsigned)r - bi * n;
            qoff[it] = bi * 128u + cp * kCW + (lane & 1) * 4;
            koff[it] = ((bi / n) * n + jj) * 128u + cp * kCW + (lane & 1) * 4;
            qc[it] = __ldg(reinterpret_cast<const float4*>(A.q + qoff[it]));
            kc[it] = __ldg(reinterpret_cast<const float4*>(A.k + koff[it]));
          }
        }
        DG_PROF(0)
        mbar_wait(&hacc_full[c], ti & 1);
        DG_PROF(1)
        const int hs = kDedIO ? 0 : (hcount & 1);
        const uint32_t hb_par = (kDedIO ? (hcount & 1) : ((hcount >> 1) & 1)) ^ 1;
        if (!kLateHb) mbar_wait(&hb_empty[hs], hb_par);      // (kLateHb: waited for below, after the accumulator read and the math)
        DG_PROF(2)
        tc_fence_after();
        uint8_t* hblk = sH + hs * kWStage + (cp >> 1) * kBlkBytes;   // the operand block (64 hidden columns) this warp's 32 belong to
        float v[32];
        tmem_ld32(tmem_base + lane_base + c * 128 + cp * kCW, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hacc_empty[c]);
        DG_PROF(3)
        if (kMode == kBwdB && use_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (mk >> i) & 1u ? v[i] : 0.f;
        } else if (kMode == kBwdB) {
          float gw[16];                                              // 32 bf16 sign masks of this row
          gather_finish(gq, 2, stg, lane, gw);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t bits = __float_as_uint(gw[i]);
            v[2 * i] = (short)(bits & 0xFFFF) > 0 ? v[2 * i] : 0.f;
            v[2 * i + 1] = (short)(bits >> 16) > 0 ? v[2 * i + 1] : 0.f;
          }
        } else if (kMode == kAttn) {
          // E = acc + be;  A = ((q_i k_j) c) (E + 1) E   (layers.py:116,123-125);  this thread's row r = ((b N + i) N + j)
          const float* bb = sB1 + cp * kCW;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = ld4(bb + i);
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
          // dedicated I/O tile: the previous tile's stores (issued >= 1.5 k cycles ago) have been read by now.  The same wait
          // (the issuer's, then the pair barrier) makes the pair's operand rows reusable.
          if (lane == 0 && ti > 0) bulk_wait_read0();
          if (spill) pair_sync(); else __syncwarp();
          if (A.e_out != nullptr) {
            // E (fp32, for the backward) leaves as a TMA tile too: own row into the I/O tile (conflict-free, no transposition),
            // one elected store per warp
#pragma unroll
            for (int j = 0; j < 8; ++j) st4(io_chunk(j), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&A.tm_e, cp * kCW, (int)wrow0, iobox);
              bulk_commit();
              bulk_wait_read0();                     // (shared memory read: the residual rows may land on top of it)
            }
          }
          // this tile's residual rows are requested here and land while the modulation below runs
          if (lane == 0) {
            mbar_expect_tx(&io_full[warp], 32 * 128);
            tma_load_2d(iobox, &A.tm_x, cp * kCW, (int)wrow0, &io_full[warp]);
          }
          __syncwarp();
          // S = c q_i k_j is formed on the COALESCED side of the staging transpose (lane = 16 bytes of one of 16 rows: the
          // k rows of consecutive j are consecutive in memory, the q row is shared by ~N rows), 8 channels at a time, the
          // next group's loads in flight; a thread-per-row read of k would touch 32 different lines per instruction
          const float cs = A.cscale;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float4 qn[2], kn[2];
            if (g < 3) {
#pragma unroll
              for (int it = 0; it < 2; ++it) {
                qn[it] = __ldg(reinterpret_cast<const float4*>(A.q + qoff[it] + (g + 1) * 8));
                kn[it] = __ldg(reinterpret_cast<const float4*>(A.k + koff[it] + (g + 1) * 8));
              }
            }
#pragma unroll
            for (int it = 0; it < 2; ++it)
              st4(stg + stg_off(it * 16 + (lane >> 1), lane & 1), make_float4(qc[it].x * kc[it].x * cs, qc[it].y * kc[it].y * cs,
                                                                             qc[it].z * kc[it].z * cs, qc[it].w * kc[it].w * cs));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float4 s4 = ld4(stg + stg_off(lane, i));
              float* vv = v + g * 8 + 4 * i;
              vv[0] = s4.x * fmaf(vv[0], vv[0], vv[0]); vv[1] = s4.y * fmaf(vv[1], vv[1], vv[1]);
              vv[2] = s4.z * fmaf(vv[2], vv[2], vv[2]); vv[3] = s4.w * fmaf(vv[3], vv[3], vv[3]);
            }
            __syncwarp();
            if (g < 3) {
#pragma unroll
              for (int it = 0; it < 2; ++it) { qc[it] = qn[it]; kc[it] = kn[it]; }
            }
          }
        } else {
          const float* bb = sB1 + c * 128 + cp * kCW;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = ld4(bb + i);
            v[i] = fmaxf(v[i] + b4.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + b4.y, 0.f);
            v[i + 2] = fmaxf(v[i + 2] + b4.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + b4.w, 0.f);
          }
          if (kMode == kBwdA && A.mask_out != nullptr && wrow0 + lane < R) {
            // the sign of h is all the dgrad chain needs of it: 4 bytes per thread instead of 64
            uint32_t m32 = 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) m32 |= (v[i] > 0.f ? 1u : 0u) << i;
            reinterpret_cast<uint32_t*>(A.mask_out)[(wrow0 + lane) * (4 * HC) + c * 4 + cp] = m32;
          }
        }
        DG_PROF(12)
        if (kLateHb) {
          // single operand buffer: the previous chunk's GEMM2 has finished reading it (it ran under this chunk's accumulator read
          // and math), and the pair's side-output store out of it has been read
          mbar_wait(&hb_empty[hs], hb_par);
          tc_fence_after();
          if (spill) {
            if (issuer && lane == 0) bulk_wait_read0();
            pair_sync();
          }
        } else if (kDedIO) {
          // (the stores of the previous tile were waited for at the top of the tile)
        } else if (c == 0 && ti > 0) {
          // the previous tile's output left through TMA out of the operand buffers (see the final epilogue): they may be
          // overwritten once every warp's store has finished READING shared memory
          if (lane == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        } else if (spill) {
          if (issuer && lane == 0) bulk_wait_read1();   // the pair's side-output store out of this buffer (two chunks ago) has been read
          pair_sync();
        }
        DG_PROF(13)
        const int j0 = (cp & 1) * 4;                    // this warp's four 16-byte chunks of the block's 128-byte rows
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_block_chunk(hblk, row, j0 + j, make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
                         make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
        DG_PROF(14)
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hb_full[hs]);
        if (kLateHb && c == 0 && lane == 0) {
          // dedicated I/O tile: this tile's residual (FWD: x, BWD_B: dz) / dout (BWD_A) box is requested HERE, two chunk phases
          // before the final epilogue needs it.  The previous tile's output store out of the box was issued a whole chunk phase
          // ago: the wait for its shared-memory read returns at once.
          if (ti > 0) bulk_wait_read0();
          mbar_expect_tx(&io_full[warp], 32 * 128);
          tma_load_2d(iobox, kMode == kBwdA ? &A.tm_dout : &A.tm_x, cp * kCW, (int)wrow0, &io_full[warp]);
        }
        if (spill) {
          // the bf16 operand rows the pair just wrote ARE its rows of the side output (h / dh / scores): TMA-store them
          // from here -- box [32 rows][64 channels]
          pair_sync();
          if (issuer && lane == 0) {
            tma_store_2d(&A.tm_spill, c * 128 + (cp >> 1) * 64, (int)wrow0, hblk + q * 32 * 128);
            bulk_commit();
          }
        }
        DG_PROF(4)
        ++hcount;
      }
      // ---- final epilogue: this thread = one row, 32 columns [cp*32, +32)
      DG_PROF(5)
      // The residual rows come in, and the output rows leave, as a TMA tile: 4 boxes [128 rows][32 channels] with the
      // 128-byte swizzle, in the dedicated I/O tile sIO (64 KB; requested after the tile's first chunk, see above) -- or, built with
      // DG_CHAIN_DEDICATED_IO=0, staged in the two operand buffers sH[0..1], which are idle from the completion of the tile's
      // last GEMM2 (z_full) until the next tile's first chunk is stored.  A thread reads and later overwrites only
      // its own row part, so the tile needs no transposition and no bank conflicts (8 consecutive rows = 8 swizzle slots).
      float a[32];
      float4 xq[kMode == kBwdA ? 8 : 1];
      if (kMode == kBwdA) gather_issue(x, wrow0, R, 128, cp * kCW, 4, lane, xq);   // residual x: in flight across the wait below
      mbar_wait(z_full, ti & 1);
      DG_PROF(7)
      tc_fence_after();
      if (spill && !kDedIO) {                        // every pair's side-output stores have finished reading the operand buffers
        if (lane == 0) { bulk_wait_read0(); mbar_arrive(sp_done); }
        mbar_wait(sp_done, ti & 1);
      }
      if (!kDedIO && lane == 0) {
        mbar_expect_tx(&io_full[warp], 32 * 128);
        const CUtensorMap* tin = kMode == kBwdA ? &A.tm_dout : &A.tm_x;   // FWD/ATTN: residual x;  BWD_B: residual dz;  BWD_A: dout
        tma_load_2d(iobox, tin, cp * kCW, (int)wrow0, &io_full[warp]);
      }
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
      if (kMode == kBwdA) {
        gather_finish(xq, 4, stg, lane, a);          // (the dout tile lands meanwhile)
        DG_PROF(6)
        float v[32];
        tmem_ld32(tmem_base + lane_base + 384 + cp * kCW, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(z_empty);
        const float* bb = sB2 + cp * kCW;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float t0 = a[i] + v[i] + bb[i], t1 = a[i + 1] + v[i + 1] + bb[i + 1];
          a[i] = t0; a[i + 1] = t1;
          s1a += t0; s1b += t1; s2a = fmaf(t0, t0, s2a); s2b = fmaf(t1, t1, s2b);
        }
      } else {
        // accumulator (+ bias) into registers while the residual tile is in flight, then the residual on top of it
        tmem_ld32(tmem_base + lane_base + 384 + cp * kCW, a);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(z_empty);
        if (kMode != kBwdB) {
          const float* bb = sB2 + cp * kCW;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = ld4(bb + i);
            a[i] += b4.x; a[i + 1] += b4.y; a[i + 2] += b4.z; a[i + 3] += b4.w;
          }
        }
        mbar_wait(&io_full[warp], ti & 1);
        DG_PROF(6)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = ld4(io_chunk(j));
          const float t0 = a[4 * j] + t.x, t1 = a[4 * j + 1] + t.y, t2 = a[4 * j + 2] + t.z, t3 = a[4 * j + 3] + t.w;
          a[4 * j] = t0; a[4 * j + 1] = t1; a[4 * j + 2] = t2; a[4 * j + 3] = t3;
          s1a += t0 + t2; s1b += t1 + t3;
          s2a = fmaf(t0, t0, fmaf(t2, t2, s2a)); s2b = fmaf(t1, t1, fmaf(t3, t3, s2b));
        }
      }
      DG_PROF(8)
      // rows -> the I/O tile (own row, in place) -> one elected thread issues the TMA store of the warp's box
      auto store_tile = [&](const CUtensorMap* tm, bool wait_read) {
#pragma unroll
        for (int j = 0; j < 8; ++j) st4(io_chunk(j), make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]));
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {                             // this warp's rows only: no CTA-wide barrier on the way out
          tma_store_2d(tm, cp * kCW, (int)wrow0, iobox);
          bulk_commit();
          if (wait_read) bulk_wait_read0();
        }
        if (wait_read) __syncwarp();
      };
      if (kMode == kBwdB) {                                           // dx = dz + dh . W1
        store_tile(&A.tm_out, false);
        DG_PROF(11)
        continue;
      }
      // per-row statistics across the four column parts: [part][row] float2.  Modes whose I/O tile aliases the operand buffers
      // pass a CTA-wide barrier before the next tile's first operand store, so one copy is enough (and BWD_A's second exchange
      // uses the second 4 KB); ATTN has no such barrier: its single exchange alternates between the two copies.
      // (BWD_A with a dedicated tile: its two exchanges per tile order the reuse of one copy each -- a warp past the second barrier
      // of tile t knows every warp has read the first copy, a warp past the first barrier of tile t + 1 that every warp has read
      // the second)
      float2* st = sStats + ((kDedIO && kMode != kBwdA) ? (int)(ti & 1) * (kParts * 128) : 0);
      st[cp * 128 + row] = make_float2(s1a + s1b, s2a + s2b);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      DG_PROF(9)
      float sum1 = 0.f, sum2 = 0.f;
#pragma unroll
      for (int p = 0; p < kParts; ++p) { const float2 o = st[p * 128 + row]; sum1 += o.x; sum2 += o.y; }
      const float mean = sum1 * (1.f / 128.f);
      const float rstd = rsqrtf(fmaxf(sum2 * (1.f / 128.f) - mean * mean, 0.f) + A.eps);
      const float* gg = sG + cp * kCW;
      if (kMode == kAttn && A.z_out != nullptr) store_tile(&A.tm_z, true);   // pre-LayerNorm sum, for the LayerNorm backward
      if (kMode == kFwd || kMode == kAttn) {
        const float* be = sBe + cp * kCW;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 g4 = ld4(gg + i), e4 = ld4(be + i);
          a[i] = (a[i] - mean) * rstd * g4.x + e4.x; a[i + 1] = (a[i + 1] - mean) * rstd * g4.y + e4.y;
          a[i + 2] = (a[i + 2] - mean) * rstd * g4.z + e4.z; a[i + 3] = (a[i + 3] - mean) * rstd * g4.w + e4.w;
        }
        store_tile(&A.tm_out, false);
        DG_PROF(11)
        continue;
      }
      // ---- BWD_A: LayerNorm backward.  xh = (z - mean) rstd;  gh = gamma * dout;
      //      dz = rstd * (gh - mean(gh) - xh * mean(gh * xh));  dgamma += dout * xh, dbeta += dout (column sums)
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = (a[i] - mean) * rstd;      // a := xh
      float sg = 0.f, sgx = 0.f;
      mbar_wait(&io_full[warp], ti & 1);                              // this warp's dout rows are in the I/O tile
      // pass 1: row sums of gh and gh*xh; the dgamma (dout*xh) and dbeta (dout) column sums go across the warp in registers
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float pg[8], pd[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 d4 = ld4(io_chunk(g * 2 + i));
          const float* xh = a + g * 8 + 4 * i;
          const float4 g4 = ld4(gg + g * 8 + 4 * i);
          const float h0 = g4.x * d4.x, h1 = g4.y * d4.y, h2 = g4.z * d4.z, h3 = g4.w * d4.w;
          sg += (h0 + h1) + (h2 + h3);
          sgx = fmaf(h0, xh[0], fmaf(h1, xh[1], fmaf(h2, xh[2], fmaf(h3, xh[3], sgx))));
          pd[4 * i] = d4.x; pd[4 * i + 1] = d4.y; pd[4 * i + 2] = d4.z; pd[4 * i + 3] = d4.w;
          pg[4 * i] = d4.x * xh[0]; pg[4 * i + 1] = d4.y * xh[1]; pg[4 * i + 2] = d4.z * xh[2]; pg[4 * i + 3] = d4.w * xh[3];
        }
        if (want_affine) {                                            // (dgrad-only passes: no dgamma / dbeta column sums)
          acc_g[g] += colsum8(pg, lane);
          acc_b[g] += colsum8(pd, lane);
        }
      }
      DG_PROF(10)
      float2* st2 = sStats + kParts * 128;
      st2[cp * 128 + row] = make_float2(sg, sgx);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      DG_PROF(9)
      float tg = 0.f, tgx = 0.f;
#pragma unroll
      for (int p = 0; p < kParts; ++p) { const float2 o = st2[p * 128 + row]; tg += o.x; tgx += o.y; }
      const float c1 = tg * (1.f / 128.f), c2 = tgx * (1.f / 128.f);
      // pass 2: dz over dout, in place (a thread reads and writes its own row only), then out through TMA
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 d4 = ld4(io_chunk(j)), g4 = ld4(gg + 4 * j);
        st4(io_chunk(j), make_float4(rstd * (g4.x * d4.x - c1 - a[4 * j] * c2), rstd * (g4.y * d4.y - c1 - a[4 * j + 1] * c2),
                                     rstd * (g4.z * d4.z - c1 - a[4 * j + 2] * c2), rstd * (g4.w * d4.w - c1 - a[4 * j + 3] * c2)));
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&A.tm_out, cp * kCW, (int)wrow0, iobox);
        bulk_commit();
      }
      DG_PROF(11)
    }
    if (lane == 0) bulk_wait0();                                      // outstanding TMA stores complete before the CTA retires
    if (kMode == kBwdA && want_affine) {
      const int col = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);      // this lane's column of a group (colsum8)
#pragma unroll
      for (int g = 0; g < 4; ++g) {                                  // fold the two low lane bits; lanes 0, 4, .., 28 flush
        float tg = acc_g[g], tb = acc_b[g];
        tg += __shfl_xor_sync(0xffffffffu, tg, 1); tg += __shfl_xor_sync(0xffffffffu, tg, 2);
        tb += __shfl_xor_sync(0xffffffffu, tb, 1); tb += __shfl_xor_sync(0xffffffffu, tb, 2);
        if ((lane & 3) == 0) {
          atomicAdd(A.dgamma + cp * kCW + g * 8 + col, tg);
          atomicAdd(A.dbeta + cp * kCW + g * 8 + col, tb);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (A.prof != nullptr && tid < 64) A.prof[(long long)blockIdx.x * 64 + tid] = sProf[tid];
}

static long long* g_chain_prof = nullptr;

// [R,128] fp32 row tensor -> tensor map with box [32 rows][32 channels], 128-byte swizzle (driver entry point, no libcuda link)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_row_tmap(CUtensorMap* tm, const float* base, long long R) {
  static TmapEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      return fail("cuTensorMapEncodeTiled is not available from this driver");
    fn = (TmapEncodeFn)p;
  }
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail("TMA row tensors must be 16-byte aligned");
  const cuuint64_t dims[2] = {128, (cuuint64_t)R}, strides[1] = {512};
  const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};       // one epilogue warp's rows of a 32-channel column group
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int make_spill_tmap(CUtensorMap* tm, const uint16_t* base, long long R, int H) {
  CUtensorMap probe;
  if (make_row_tmap(&probe, reinterpret_cast<const float*>(base), 1)) return 1;    // resolves the entry point / checks alignment
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  const cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)R}, strides[1] = {(cuuint64_t)H * 2};
  const cuuint32_t box[2] = {64, 32}, estr[2] = {1, 1};
  CUresult r = ((TmapEncodeFn)p)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(bf16 side output) failed (%d)", (int)r);
  return 0;
}

template <int kMode>
static int launch_chain(const MlpArgs& a, cudaStream_t s) {
  {   // per-device attribute: set on every launch so that every device of the process is configured
    cudaError_t e = cudaFuncSetAttribute(mlp_chain_tc_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpSmem::total + 1024);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute(mlp_chain): %s", cudaGetErrorString(e));
  }
  long long tiles = (a.R + 127) / 128;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  MlpArgs b = a;
  if (make_row_tmap(&b.tm_x, a.x, a.R)) return 1;
  if (make_row_tmap(&b.tm_out, a.out, a.R)) return 1;
  if (kMode == kBwdA && make_row_tmap(&b.tm_dout, a.dout, a.R)) return 1;
  if (kMode == kAttn && a.z_out != nullptr && make_row_tmap(&b.tm_z, a.z_out, a.R)) return 1;
  if (kMode == kAttn && a.e_out != nullptr && make_row_tmap(&b.tm_e, a.e_out, a.R)) return 1;
  if (a.spill != nullptr && make_spill_tmap(&b.tm_spill, a.spill, a.R, a.HC * 128)) return 1;
  b.prefetch = (opt_get(DG_OPT_L2_PREFETCH) & DG_PF_CHAIN) | ((opt_get(DG_OPT_L2_PREFETCH) & DG_PF_CHAIN_KEEP) ? 2 : 0);
  b.prof = g_chain_prof;
  mlp_chain_tc_kernel<kMode><<<grid, kMlpThreads, MlpSmem::total + 1024, s>>>(b);
  return check_launch("dg_mlp_chain");
}

}  // namespace tc
}  // namespace dg

using namespace dg;

static int mlp_check(const char* who, const void* x, long long R, int D, int H, const void* ws, long long ws_bytes) {
  if (R <= 0) return fail("%s: rows must be > 0", who);
  if (reinterpret_cast<uintptr_t>(x) & 31) return fail("%s: the row input must be 32-byte aligned (256-bit loads)", who);
  if (D != 128 || H % 128 || H < 128 || H > 384) return fail("%s: needs D == 128 and H in {128,256,384}, got D=%d H=%d", who, D, H);
  if (ws_bytes < (long long)2 * (H / 128) * tc::kWStage) return fail("%s: workspace too small (%lld bytes)", who, ws_bytes);
  if (reinterpret_cast<uintptr_t>(ws) & 127) return fail("%s: workspace must be 128-byte aligned", who);
  return 0;
}

extern "C" int dg_mlp_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* gamma, const float* beta, float* out, long long R, int D, int H, float eps,
                          void* workspace, long long workspace_bytes, void* stream) {
  DG_TRACE("dg_mlp_fwd", x, w1, b1, w2, b2, gamma, beta, out, R, D, H, eps, workspace);
  if (mlp_check("dg_mlp_fwd", x, R, D, H, workspace, workspace_bytes)) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  tc::mlp_pack_weights_kernel<<<48, 256, 0, s>>>(w1, w2, (uint8_t*)workspace, H, 0);
  tc::MlpArgs a{x, (const uint8_t*)workspace, b1, b2, gamma, beta, nullptr, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, R, H / 128, eps};
  return tc::launch_chain<tc::kFwd>(a, s);
}

extern "C" int dg_mlp_bwd_ln(const float* x, const float* dout, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* gamma, float* dz, void* h_bf16, void* relu_mask, float* dgamma,
                             float* dbeta, long long R, int D, int H, float eps, void* workspace, long long workspace_bytes,
                             void* stream) {
  DG_TRACE("dg_mlp_bwd_ln", x, dout, w1, b1, w2, b2, gamma, dz, h_bf16, relu_mask, dgamma, dbeta, R, D, H, eps, workspace);
  if (mlp_check("dg_mlp_bwd_ln", x, R, D, H, workspace, workspace_bytes)) return 1;
  if (reinterpret_cast<uintptr_t>(relu_mask) & 7) return fail("dg_mlp_bwd_ln: the sign mask must be 8-byte aligned");
  if ((dgamma == nullptr) != (dbeta == nullptr)) return fail("dg_mlp_bwd_ln: pass both dgamma and dbeta or neither");
  cudaStream_t s = (cudaStream_t)stream;
  tc::mlp_pack_weights_kernel<<<48, 256, 0, s>>>(w1, w2, (uint8_t*)workspace, H, 0);
  tc::MlpArgs a{x, (const uint8_t*)workspace, b1, b2, gamma, nullptr, dout, dz, (uint16_t*)h_bf16, nullptr,
                (unsigned long long*)relu_mask, nullptr, dgamma, dbeta, R, H / 128, eps};
  return tc::launch_chain<tc::kBwdA>(a, s);
}

extern "C" int dg_mlp_bwd_dgrad(const float* dz, const void* h_bf16, const void* relu_mask, const float* w1, const float* w2,
                                float* dx, void* dh_bf16, long long R, int D, int H, void* workspace, long long workspace_bytes,
                                void* stream) {
  DG_TRACE("dg_mlp_bwd_dgrad", dz, h_bf16, relu_mask, w1, w2, dx, dh_bf16, R, D, H, workspace);
  if (mlp_check("dg_mlp_bwd_dgrad", dz, R, D, H, workspace, workspace_bytes)) return 1;
  if (h_bf16 == nullptr && relu_mask == nullptr) return fail("dg_mlp_bwd_dgrad: needs h (bf16) or its sign mask");
  if (reinterpret_cast<uintptr_t>(relu_mask) & 7) return fail("dg_mlp_bwd_dgrad: the sign mask must be 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  tc::mlp_pack_weights_kernel<<<48, 256, 0, s>>>(w1, w2, (uint8_t*)workspace, H, 1);
  tc::MlpArgs a{dz, (const uint8_t*)workspace, nullptr, nullptr, nullptr, nullptr, nullptr, dx, (uint16_t*)dh_bf16,
                (const uint16_t*)h_bf16, nullptr, (const unsigned long long*)relu_mask, nullptr, nullptr, R, H / 128, 0.f};
  return tc::launch_chain<tc::kBwdB>(a, s);
}

extern "C" int dg_attn_edge_fwd(const float* y, const float* q, const float* k, const float* we, const float* be,
                                const float* woe, const float* boe, const float* gamma, const float* beta, float c,
                                float* out, void* a_bf16, float* e_out, float* z_out, int B, int N, int D, float eps,
                                void* workspace, long long workspace_bytes, void* stream) {
  DG_TRACE("dg_attn_edge_fwd", y, q, k, we, be, woe, boe, gamma, beta, c, out, a_bf16, e_out, z_out, B, N, D, eps, workspace);
  if (B <= 0 || N <= 0) return fail("dg_attn_edge_fwd: bad shape B=%d N=%d", B, N);
  if ((long long)B * N * N >= (1ll << 31) || (long long)B * N * 128 >= (1ll << 31)) return fail("dg_attn_edge_fwd: B*N*N and B*N*128 must be < 2^31 (split the batch)");
  const long long R = (long long)B * N * N;
  if (mlp_check("dg_attn_edge_fwd", y, R, D, 128, workspace, workspace_bytes)) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  tc::mlp_pack_weights_kernel<<<48, 256, 0, s>>>(we, woe, (uint8_t*)workspace, 128, 0);
  tc::MlpArgs a{y, (const uint8_t*)workspace, be, boe, gamma, beta, nullptr, out, (uint16_t*)a_bf16, nullptr, nullptr, nullptr, nullptr, nullptr,
                R, 1, eps, q, k, e_out, z_out, N, c, 0};
  return tc::launch_chain<tc::kAttn>(a, s);
}

/* debug: phase cycle counters of the chain kernels ([148 CTAs][4 roles][16] int64 device buffer, or NULL to stop) */
extern "C" int dg_debug_chain_profile(void* device_buf) {
  tc::g_chain_prof = (long long*)device_buf;
  return 0;
}

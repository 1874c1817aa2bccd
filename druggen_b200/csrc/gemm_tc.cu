// tcgen05 contractions (DG_PREC_BF16): bf16 operands, fp32 accumulation in TMEM.
//
//   rows_gemm_tc : out[R,N] = epi(a[R,K] . op(w) + bias)   K,N in {64..384}, multiples of 64 / 128
//   gemm_tn_tc   : out[M,N] += a[R,M]^T . b[R,N]           split over rows, atomics at the end
//
// Both are persistent, warp-specialised kernels (one CTA per SM):
//   epilogue warps : tcgen05.ld (thread = accumulator row) -> smem transpose -> coalesced stores
//   loader warps   : coalesced fp32 LDG.128 -> bf16 -> 128B-swizzled operand blocks in smem
//   MMA warp       : one elected thread issues tcgen05.mma, tcgen05.commit signals mbarriers
// Two loader groups (alternate ring stages) and, in rows_gemm, two epilogue groups (alternate
// accumulator buffers) keep >= 64 KB of loads in flight per SM: with ~2 us of loaded HBM latency a
// single 4-warp group caps the kernel at ~4 TB/s.
// Activations are fp32 in HBM in this (unfused) form, so both kernels are HBM-bound: the roofline
// that governs them is bytes moved (a + out, or a + b), not the tensor pipe.
#include "tc_common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {
namespace tc {

constexpr int kThreads = 544;            // rows_gemm: warps 0-7 epilogue, 8-15 loaders, 16 MMA
constexpr int kTnThreads = 416;          // gemm_tn : warps 0-3 epilogue, 4-11 loaders, 12 MMA
constexpr int kBlk = 128 * 128;           // [128 rows][64 bf16] operand block, bytes
constexpr int kStage = 36;                // epilogue transpose row pitch (floats): conflict-free v4 access

// =============================================================================================
// rows_gemm
// =============================================================================================
constexpr int kRingA = 4;                 // A ring: blocks of [128 rows][64 ch]
constexpr int kAccBufs = 4;               // 4 x 128 TMEM columns
constexpr int kResidPre = 256;            // internal flag: resid is added before the ReLU / gate (K-sliced launches)

struct RowsSmem {                         // offsets into dynamic smem (1024-aligned base)
  int w, a, stage, bars, total;
};
// split (DG_PREC_BF16X3): every operand block exists twice (hi, lo); the A ring is then 3 deep so that it all still fits
__host__ __device__ inline int rows_ring(int split) { return split ? 3 : kRingA; }
__host__ __device__ inline RowsSmem rows_smem(int K, int N, int split = 0) {
  RowsSmem s;
  s.w = 0;
  s.a = N * K * 2 * (split ? 2 : 1);
  s.stage = s.a + rows_ring(split) * kBlk * (split ? 2 : 1);
  s.bars = s.stage + 8 * 32 * kStage * 4;
  s.total = s.bars + 256;
  return s;
}

__global__ void __launch_bounds__(kThreads, 1)
rows_gemm_tc_kernel(const float* __restrict__ a, const float* __restrict__ w, int w_is_nk,
                    const float* __restrict__ bias, int relu, const float* __restrict__ gate,
                    const float* __restrict__ resid, float* __restrict__ out, long long R, int K, int N, int flags,
                    int prefetch, int lda, int ldw, int ldo, int split) {
  // lda / ldw / ldo: row pitches (elements) of a, w and of out / resid / gate -- a launch may work on a column slice of wider
  // tensors (the split-precision mode runs H = 384 shapes as 128-wide slices).  split: bf16x3 operands (hi + lo blocks).
  const uint16_t* a16 = reinterpret_cast<const uint16_t*>(a);        // bf16 views (flags select which are live)
  const uint16_t* gate16 = reinterpret_cast<const uint16_t*>(gate);
  uint16_t* out16 = reinterpret_cast<uint16_t*>(out);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: a pointer rebuilt from an integer loses its address
  // space and every access through it compiles to generic LD/ST (LSU long-scoreboard path) instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const RowsSmem L = rows_smem(K, N, split);
  const int ring = rows_ring(split);
  const int a_stage = split ? 2 * kBlk : kBlk;          // ring stage: [hi block][lo block]
  const int w_lo = N * K * 2;                           // offset of the lo copy of the weights
  uint8_t* sW = smem + L.w;
  uint8_t* sA = smem + L.a;
  float* sStage = reinterpret_cast<float*>(smem + L.stage);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* a_empty = a_full + kRingA;                   // (barrier slots are laid out for the deepest ring)
  uint64_t* acc_full = a_empty + kRingA;
  uint64_t* acc_empty = acc_full + kAccBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccBufs);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = K / 64, NC = N / 128;
  const long long num_tiles = (R + 127) / 128;

  if (tid == 0) {
    for (int i = 0; i < kRingA; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kAccBufs; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  // weights: fp32 (L2-resident) -> bf16 -> [kb][N rows][128 B] swizzled, resident for the whole kernel
  for (int idx = tid; idx < N * (K / 8); idx += kThreads) {
    int n, c8;
    float4 lo, hi;
    if (w_is_nk) {
      n = idx / (K / 8); c8 = idx % (K / 8);
      const float* p = w + (long long)n * ldw + c8 * 8;
      lo = ld4(p); hi = ld4(p + 4);
    } else {
      n = idx % N; c8 = idx / N;
      const float* p = w + (long long)(c8 * 8) * ldw + n;
      lo = make_float4(p[0], p[ldw], p[2 * ldw], p[3 * ldw]);
      hi = make_float4(p[4 * ldw], p[5 * ldw], p[6 * ldw], p[7 * ldw]);
    }
    if (split) st_block_chunk_split(sW + (c8 >> 3) * (N * 128), sW + w_lo + (c8 >> 3) * (N * 128), n, c8 & 7, lo, hi);
    else st_block_chunk(sW + (c8 >> 3) * (N * 128), n, c8 & 7, lo, hi);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8 && warp < 16) {
    // ------------------------------------------------------------------ loaders (two groups, alternate chunks)
    const int lt = (tid - 256) & 127, grp = (tid - 256) >> 7;
    uint32_t chunk = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long row0 = tile * 128;
      if (prefetch && lt == 0) {
        // TMA-engine L2 prefetch of whole contiguous row tiles: group 0 -> the operand rows two tiles ahead,
        // group 1 -> what the epilogue reads (resid / gate) one tile ahead
        const long long first = (tile == (long long)blockIdx.x) ? 1 : (grp == 0 ? 2 : 1), last = grp == 0 ? 2 : 1;
        for (long long ahead = first; ahead <= last; ++ahead) {
          const long long prow0 = (tile + ahead * gridDim.x) * 128;
          if (prow0 >= R) break;
          const long long prows = R - prow0 < 128 ? R - prow0 : 128;
          if (grp == 0) {
            if (lda == K) bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(a) + prow0 * K * ((flags & DG_A_BF16) ? 2 : 4), prows * K * ((flags & DG_A_BF16) ? 2 : 4));
          } else if (ldo == N) {
            if (resid) bulk_prefetch_l2(resid + prow0 * N, prows * N * 4);
            if (gate) bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(gate) + prow0 * N * ((flags & DG_GATE_BF16) ? 2 : 4), prows * N * ((flags & DG_GATE_BF16) ? 2 : 4));
          }
        }
      }
      for (int kb = 0; kb < KB; ++kb, ++chunk) {
        if ((int)(chunk & 1) != grp) continue;
        const int st = chunk % ring;
        mbar_wait(&a_empty[st], ((chunk / ring) & 1) ^ 1);
        uint8_t* blk = sA + st * a_stage;
        if (flags & DG_A_BF16) {            // operand already bf16 in HBM: plain 16-byte chunk copies
          uint4 c16[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            int item = it * 128 + lt, r = item >> 3, j = item & 7;
            c16[it] = (row0 + r < R) ? *reinterpret_cast<const uint4*>(a16 + (row0 + r) * lda + kb * 64 + j * 8) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            int item = it * 128 + lt, r = item >> 3, j = item & 7;
            *reinterpret_cast<uint4*>(blk + r * 128 + ((j ^ (r & 7)) << 4)) = c16[it];
          }
          fence_async_smem();
          mbar_arrive(&a_full[st]);
          continue;
        }
        float4 v[16];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          int item = it * 128 + lt, r = item >> 3, j = item & 7;
          if (row0 + r < R) {
            ld8(a + (row0 + r) * lda + kb * 64 + j * 8, v[2 * it], v[2 * it + 1]);
          } else {
            v[2 * it] = v[2 * it + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          int item = it * 128 + lt;
          if (split) st_block_chunk_split(blk, blk + kBlk, item >> 3, item & 7, v[2 * it], v[2 * it + 1]);
          else st_block_chunk(blk, item >> 3, item & 7, v[2 * it], v[2 * it + 1]);
        }
        fence_async_smem();
        mbar_arrive(&a_full[st]);
      }
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, 128, 0, 0);
      uint32_t chunk0 = 0, unit = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, chunk0 += KB) {
        for (int nc = 0; nc < NC; ++nc, ++unit) {
          const int buf = unit % kAccBufs;
          mbar_wait(&acc_empty[buf], ((unit / kAccBufs) & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t chunk = chunk0 + kb;
            const int st = chunk % ring;
            if (nc == 0) {
              mbar_wait(&a_full[st], (chunk / ring) & 1);
              tc_fence_after();
            }
            const uint32_t a_addr = smem_u32(sA + st * a_stage);
            const uint32_t b_addr = smem_u32(sW + kb * (N * 128) + nc * kBlk);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tmem_base + buf * 128, make_sdesc(a_addr + k * 32, 16, 1024), make_sdesc(b_addr + k * 32, 16, 1024),
                        idesc, (kb | k) ? 1u : 0u);
              if (split) {       // + a_hi . w_lo + a_lo . w_hi into the same fp32 accumulator
                umma_bf16(tmem_base + buf * 128, make_sdesc(a_addr + k * 32, 16, 1024), make_sdesc(b_addr + w_lo + k * 32, 16, 1024), idesc, 1u);
                umma_bf16(tmem_base + buf * 128, make_sdesc(a_addr + kBlk + k * 32, 16, 1024), make_sdesc(b_addr + k * 32, 16, 1024), idesc, 1u);
              }
            }
            if (nc == NC - 1) umma_commit(&a_empty[st]);
          }
          umma_commit(&acc_full[buf]);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-7: two groups, alternate units)
    float* stg = sStage + warp * 32 * kStage;
    const int qw = warp & 3, egrp = warp >> 2;
    uint32_t unit = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long row0 = tile * 128 + qw * 32;
      for (int nc = 0; nc < NC; ++nc, ++unit) {
        if ((int)(unit & 1) != egrp) continue;
        const int buf = unit % kAccBufs;
        if (flags & DG_OUT_BF16) {
          // bf16 output (+ optional bf16 sign gate): lane = (row it*8 + lane/4, 8 columns); the whole unit's gate
          // (32 rows x 128 cols) is requested before the accumulator is even waited for
          const int c8 = (lane & 3) * 8;
          uint4 gq[8];                      // ring over column groups: cg and cg+1 in flight
          auto gate_fetch = [&](int cg) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const long long grow = row0 + it * 8 + (lane >> 2);
              gq[(cg & 1) * 4 + it] = (grow < R) ? *reinterpret_cast<const uint4*>(gate16 + grow * ldo + nc * 128 + cg * 32 + c8)
                                                 : make_uint4(0, 0, 0, 0);
            }
          };
          if (gate) { gate_fetch(0); gate_fetch(1); }
          mbar_wait(&acc_full[buf], (unit / kAccBufs) & 1);
          tc_fence_after();
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(qw * 32) << 16) + buf * 128 + cg * 32, v);
            tmem_ld_wait();
            if (cg == 3) { tc_fence_before(); mbar_arrive(&acc_empty[buf]); }
#pragma unroll
            for (int i = 0; i < 8; ++i) st4(stg + lane * kStage + i * 4, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            __syncwarp();
            const int col = nc * 128 + cg * 32 + c8;
            const float4 b0 = bias ? ld4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = bias ? ld4(bias + col + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + (lane >> 2);
              const long long grow = row0 + r;
              float4 lo = ld4(stg + r * kStage + c8), hi = ld4(stg + r * kStage + c8 + 4);
              float f[8] = {lo.x + b0.x, lo.y + b0.y, lo.z + b0.z, lo.w + b0.w, hi.x + b1.x, hi.y + b1.y, hi.z + b1.z, hi.w + b1.w};
              if (relu) {
#pragma unroll
                for (int e8 = 0; e8 < 8; ++e8) f[e8] = fmaxf(f[e8], 0.f);
              }
              if (gate) {       // bf16 > 0  <=>  its 16 bits, read as a signed short, are > 0
                const uint4 gb = gq[(cg & 1) * 4 + it];
                const uint32_t gw[4] = {gb.x, gb.y, gb.z, gb.w};
#pragma unroll
                for (int e8 = 0; e8 < 8; ++e8) {
                  const short bits = (short)((gw[e8 >> 1] >> ((e8 & 1) * 16)) & 0xFFFF);
                  f[e8] = bits > 0 ? f[e8] : 0.f;
                }
              }
              if (grow < R)
                *reinterpret_cast<uint4*>(out16 + grow * ldo + col) =
                    make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
            if (gate && cg + 2 < 4) gate_fetch(cg + 2);
            __syncwarp();
          }
          continue;
        }
        mbar_wait(&acc_full[buf], (unit / kAccBufs) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cg = 0; cg < 4; ++cg) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(qw * 32) << 16) + buf * 128 + cg * 32, v);
          tmem_ld_wait();
          if (cg == 3) {                      // accumulator fully read: hand the TMEM buffer back
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) st4(stg + lane * kStage + i * 4, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
          __syncwarp();
          const int col = nc * 128 + cg * 32 + (lane & 7) * 4;
          const float4 bz = bias ? ld4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          // two batches of 4 rows: all global loads of a batch (gate, resid) are issued before any is used
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float4 o[4], gz[4], rz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = (half * 4 + u) * 4 + (lane >> 3);
              const long long grow = row0 + r;
              o[u] = ld4(stg + r * kStage + (lane & 7) * 4);
              gz[u] = (gate && grow < R) ? ld4(gate + grow * ldo + col) : make_float4(1.f, 1.f, 1.f, 1.f);
              rz[u] = (resid && grow < R) ? ld4(resid + grow * ldo + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = (half * 4 + u) * 4 + (lane >> 3);
              const long long grow = row0 + r;
              float4 v4 = o[u];
              v4.x += bz.x; v4.y += bz.y; v4.z += bz.z; v4.w += bz.w;
              if (flags & kResidPre) {     // resid = the partial sum of earlier K slices: it belongs INSIDE the ReLU / gate
                v4.x += rz[u].x; v4.y += rz[u].y; v4.z += rz[u].z; v4.w += rz[u].w;
                rz[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              if (relu) { v4.x = fmaxf(v4.x, 0.f); v4.y = fmaxf(v4.y, 0.f); v4.z = fmaxf(v4.z, 0.f); v4.w = fmaxf(v4.w, 0.f); }
              v4.x = (gz[u].x > 0.f ? v4.x : 0.f) + rz[u].x;
              v4.y = (gz[u].y > 0.f ? v4.y : 0.f) + rz[u].y;
              v4.z = (gz[u].z > 0.f ? v4.z : 0.f) + rz[u].z;
              v4.w = (gz[u].w > 0.f ? v4.w : 0.f) + rz[u].w;
              if (grow < R) st4(out + grow * ldo + col, v4);
            }
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =============================================================================================
// gemm_tn : out[M,N] += a[R,M]^T b[R,N]
// =============================================================================================
constexpr int kTnRows = 64;               // rows (= MMA K extent) per stage
constexpr int kTnBlk = kTnRows * 128;     // [64 rows][64 ch] block, bytes

__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_tc_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                  float* __restrict__ colsum_a, long long R, int M, int N, long long tiles_per_cta, int stages, int flags,
                  int prefetch, int lda, int ldb, int ldo, int split) {
  // lda / ldb / ldo: row pitches (elements) of a, b, out (column slices of wider tensors); split: bf16x3 operands -- every
  // stage holds the hi blocks of a and b followed by their lo blocks
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: a pointer rebuilt from an integer loses its address
  // space and every access through it compiles to generic LD/ST (LSU long-scoreboard path) instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nblk = (M + N) / 64;                       // operand blocks per stage: a's first, then b's
  const int lo_off = nblk * kTnBlk;                    // split: offset of the lo copies inside a stage
  const int stage_bytes = nblk * kTnBlk * (split ? 2 : 1);
  uint8_t* sOp = smem;
  float* sStage = reinterpret_cast<float*>(smem + stages * stage_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes + 4 * 32 * kStage * 4);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long num_tiles = (R + kTnRows - 1) / kTnRows;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta;
  const long long t1 = t0 + tiles_per_cta < num_tiles ? t0 + tiles_per_cta : num_tiles;
  const int MB = M / 128;

  if (tid == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 12) {
    const int lt = (tid - 128) & 127, grp = (tid - 128) >> 7;     // two loader groups, alternate stages
    uint32_t it_ = 0;
    // bias gradient on the side: this thread always loads the same 8 channels (j = lt & 7) of every a-block,
    // so exact fp32 column sums of `a` accumulate in registers for free
    float cs[6][8];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) cs[i][e] = 0.f;
    for (long long tile = t0; tile < t1; ++tile, ++it_) {
      if ((int)(it_ & 1) != grp) continue;
      if (prefetch && lt == 0) {        // L2 prefetch (TMA engine) of this group's tile after next (contiguous a / b rows)
        const long long p0 = (tile + (it_ < 2 ? 1 : 4)) * kTnRows, p1 = (tile + 5 < t1 ? tile + 5 : t1) * kTnRows;
        const long long pe = p1 < R ? p1 : R;
        if (pe > p0) {
          if (lda == M) bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(a) + p0 * M * ((flags & DG_A_BF16) ? 2 : 4), (pe - p0) * M * ((flags & DG_A_BF16) ? 2 : 4));
          if (ldb == N) bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(b) + p0 * N * ((flags & DG_OUT_BF16) ? 2 : 4), (pe - p0) * N * ((flags & DG_OUT_BF16) ? 2 : 4));
        }
      }
      const int st = it_ % stages;
      mbar_wait(&empty[st], ((it_ / stages) & 1) ^ 1);
      uint8_t* base = sOp + st * stage_bytes;
      const long long row0 = tile * kTnRows;
      // items: (block, row, chunk j); 64 rows x 8 chunks = 512 items per block, 4 per thread
      // Blocks are loaded in batches that keep 16 independent 16-byte loads in flight per thread:
      // two fp32 blocks (8 x float4 each) or four bf16 blocks (4 x uint4 each, copied to smem without conversion).
      const int MA = M / 64;
#pragma unroll
      for (int seg = 0; seg < 2; ++seg) {
        const float* src = seg == 0 ? a : b;
        const int ld = seg == 0 ? lda : ldb, nb = seg == 0 ? MA : nblk - MA, blk_off = seg == 0 ? 0 : MA;
        const bool src16 = seg == 0 ? (flags & DG_A_BF16) : (flags & DG_OUT_BF16);
        const bool sum = seg == 0 && colsum_a != nullptr;
        if (src16) {
          const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
#pragma unroll
          for (int g0 = 0; g0 < 6; g0 += 4) {
            if (g0 >= nb) break;
            uint4 c16[16];
#pragma unroll
            for (int h = 0; h < 4; ++h)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                int item = q * 128 + lt, r = item >> 3, j = item & 7;
                c16[h * 4 + q] = (g0 + h < nb && row0 + r < R) ? *reinterpret_cast<const uint4*>(s16 + (row0 + r) * ld + (g0 + h) * 64 + j * 8)
                                                                : make_uint4(0, 0, 0, 0);
              }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              if (g0 + h >= nb) break;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                int item = q * 128 + lt, r = item >> 3, j = item & 7;
                *reinterpret_cast<uint4*>(base + (blk_off + g0 + h) * kTnBlk + r * 128 + ((j ^ (r & 7)) << 4)) = c16[h * 4 + q];
                if (sum && g0 + h < 6) {
                  const uint4 c = c16[h * 4 + q];
                  float* acc = cs[g0 + h];
                  acc[0] += __uint_as_float(c.x << 16); acc[1] += __uint_as_float(c.x & 0xFFFF0000u);
                  acc[2] += __uint_as_float(c.y << 16); acc[3] += __uint_as_float(c.y & 0xFFFF0000u);
                  acc[4] += __uint_as_float(c.z << 16); acc[5] += __uint_as_float(c.z & 0xFFFF0000u);
                  acc[6] += __uint_as_float(c.w << 16); acc[7] += __uint_as_float(c.w & 0xFFFF0000u);
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int g0 = 0; g0 < 6; g0 += 2) {
            if (g0 >= nb) break;
            float4 v[16];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                int item = q * 128 + lt, r = item >> 3, j = item & 7;
                if (row0 + r < R) {
                  ld8(src + (row0 + r) * ld + (g0 + h) * 64 + j * 8, v[h * 8 + 2 * q], v[h * 8 + 2 * q + 1]);
                } else {
                  v[h * 8 + 2 * q] = v[h * 8 + 2 * q + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                int item = q * 128 + lt;
                uint8_t* bp = base + (blk_off + g0 + h) * kTnBlk;
                if (split) st_block_chunk_split(bp, bp + lo_off, item >> 3, item & 7, v[h * 8 + 2 * q], v[h * 8 + 2 * q + 1]);
                else st_block_chunk(bp, item >> 3, item & 7, v[h * 8 + 2 * q], v[h * 8 + 2 * q + 1]);
              }
              if (sum && g0 + h < 6) {
                float* acc = cs[g0 + h];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  acc[0] += v[h * 8 + 2 * q].x; acc[1] += v[h * 8 + 2 * q].y; acc[2] += v[h * 8 + 2 * q].z; acc[3] += v[h * 8 + 2 * q].w;
                  acc[4] += v[h * 8 + 2 * q + 1].x; acc[5] += v[h * 8 + 2 * q + 1].y; acc[6] += v[h * 8 + 2 * q + 1].z; acc[7] += v[h * 8 + 2 * q + 1].w;
                }
              }
            }
          }
        }
      }
      fence_async_smem();
      mbar_arrive(&full[st]);
    }
    if (colsum_a != nullptr) {
#pragma unroll
      for (int blk = 0; blk < 6; ++blk) {
        if (blk >= M / 64) break;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float s = cs[blk][e];
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          if (lane < 8 && s != 0.f) atomicAdd(colsum_a + blk * 64 + lane * 8 + e, s);
        }
      }
    }
  } else if (warp == 12) {
    if (lane == 0) {
      uint32_t it_ = 0;
      for (long long tile = t0; tile < t1; ++tile, ++it_) {
        const int st = it_ % stages;
        mbar_wait(&full[st], (it_ / stages) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sOp + st * stage_bytes);
        const uint32_t b_addr = a_addr + (M / 64) * kTnBlk;
#pragma unroll 1
        for (int ks = 0; ks < kTnRows / 16; ++ks) {
          for (int mb = 0; mb < MB; ++mb) {
            // A: MN-major, M = 128 channels = 2 blocks (LBO = block pitch), K = 16 rows = 2 atoms (SBO = 1024)
            const uint64_t da = make_sdesc(a_addr + mb * 2 * kTnBlk + ks * 2048, kTnBlk, 1024);
            for (int n0 = 0; n0 < N; n0 += 256) {
              const int nn = N - n0 < 256 ? N - n0 : 256;
              const uint64_t db = make_sdesc(b_addr + (n0 / 64) * kTnBlk + ks * 2048, kTnBlk, 1024);
              umma_bf16(tmem_base + mb * N + n0, da, db, make_idesc(128, nn, 1, 1), (it_ | ks) ? 1u : 0u);
              if (split) {       // + a_hi^T b_lo + a_lo^T b_hi into the same fp32 accumulator
                umma_bf16(tmem_base + mb * N + n0, da, make_sdesc(b_addr + lo_off + (n0 / 64) * kTnBlk + ks * 2048, kTnBlk, 1024),
                          make_idesc(128, nn, 1, 1), 1u);
                umma_bf16(tmem_base + mb * N + n0, make_sdesc(a_addr + lo_off + mb * 2 * kTnBlk + ks * 2048, kTnBlk, 1024), db,
                          make_idesc(128, nn, 1, 1), 1u);
              }
            }
          }
        }
        umma_commit(&empty[st]);
      }
      umma_commit(done);
    }
    __syncwarp();
  } else {
    // epilogue: once, after the last tile of this CTA
    if (t1 > t0) {
      mbar_wait(done, 0);
      tc_fence_after();
      float* stg = sStage + warp * 32 * kStage;
      for (int mb = 0; mb < MB; ++mb) {
        for (int cg = 0; cg < N / 32; ++cg) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * N + cg * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) st4(stg + lane * kStage + i * 4, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int r = rr * 4 + (lane >> 3);
            const int m = mb * 128 + warp * 32 + r;
            float4 o = ld4(stg + r * kStage + (lane & 7) * 4);
            float* dst = out + (long long)m * ldo + cg * 32 + (lane & 7) * 4;
            atomicAdd(dst + 0, o.x); atomicAdd(dst + 1, o.y); atomicAdd(dst + 2, o.z); atomicAdd(dst + 3, o.w);
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc

int rows_gemm_fp32(const float*, const float*, int, const float*, int, const float*, const float*, float*, long long, int, int, cudaStream_t);
int gemm_tn_fp32(const float*, const float*, float*, float*, long long, int, int, cudaStream_t);

// Shapes the tensor-core kernels take; anything else (tiny / odd node-side shapes) runs the fp32
// CUDA-core kernel -- still on the GPU, never a CPU path.
static bool rows_tc_ok(int K, int N) {
  if (K % 64 || N % 128 || K < 64 || N < 128 || K > 384 || N > 384) return false;
  if (N > 128 && K / 64 > tc::kRingA) return false;        // a row tile must stay resident across N chunks
  return tc::rows_smem(K, N).total + 1024 <= 227 * 1024;
}

static int rows_gemm_tc_launch(const float* a, const float* w, int w_is_nk, const float* bias, int relu, const float* gate,
                               const float* resid, float* out, long long R, int K, int N, int flags, int lda, int ldw, int ldo,
                               int split, cudaStream_t s) {
  const int smem = tc::rows_smem(K, N, split).total + 1024;
  {   // per-device attribute: set on every launch (a host-side table lookup) so that every device of the process is configured
    cudaError_t e = cudaFuncSetAttribute(tc::rows_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute(rows_gemm_tc): %s", cudaGetErrorString(e));
  }
  long long tiles = (R + 127) / 128;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  tc::rows_gemm_tc_kernel<<<grid, tc::kThreads, smem, s>>>(a, w, w_is_nk, bias, relu, gate, resid, out, R, K, N, flags,
                                                           opt_get(DG_OPT_L2_PREFETCH) & DG_PF_ROWS_GEMM, lda, ldw, ldo, split);
  return check_launch(split ? "dg_rows_gemm(bf16x3)" : "dg_rows_gemm(bf16)");
}

int rows_gemm_tc(const void* a_, const float* w, int w_is_nk, const float* bias, int relu, const void* gate_,
                 const float* resid, void* out_, long long R, int K, int N, int prec, int flags, cudaStream_t s) {
  const float* a = (const float*)a_;
  const float* gate = (const float*)gate_;
  float* out = (float*)out_;
  const bool split = prec == DG_PREC_BF16X3;
  // shapes wider than one launch takes (mlp_ratio = 4: H = 512) run as 128-wide column slices, like the split-precision mode
  const bool sliced = split || (!rows_tc_ok(K, N) && K % 128 == 0 && N % 128 == 0 && K >= 128 && N >= 128 && !flags);
  if ((!sliced && !rows_tc_ok(K, N)) || (split && (K % 128 || N % 128))) {
    if (flags) return fail("dg_rows_gemm: bf16 storage is only available for the tcgen05 shapes (K=%d N=%d)", K, N);
    return rows_gemm_fp32(a, w, w_is_nk, bias, relu, gate, resid, out, R, K, N, s);
  }
  if (!(flags & DG_A_BF16) && (reinterpret_cast<uintptr_t>(a) & 31)) return fail("dg_rows_gemm: a must be 32-byte aligned (256-bit loads)");
  if (sliced) {
    // Split-precision parity mode: three bf16 MMAs per product term on hi / lo operand blocks, fp32 everywhere else.  Twice the
    // shared memory per operand, so a launch takes a 128 x 128 weight block: wider shapes run as column slices -- N slices are
    // independent, K slices accumulate through the resid path (the epilogue must then be linear: no ReLU / gate on K > 128).
    if (flags) return fail("dg_rows_gemm: the bf16x3 mode keeps every tensor fp32 (no bf16 storage flags)");
    const bool nonlinear = relu || gate;
    if (K > 128 && nonlinear && resid) return fail("dg_rows_gemm(sliced): resid together with a ReLU / gate epilogue needs K == 128, got K=%d", K);
    const int ldw = w_is_nk ? K : N;
    for (int n0 = 0; n0 < N; n0 += 128)
      for (int k0 = 0; k0 < K; k0 += 128) {
        const bool first = k0 == 0, last = k0 + 128 >= K;
        const float* ws = w_is_nk ? w + (long long)n0 * K + k0 : w + (long long)k0 * N + n0;
        // K slices accumulate through the resid path: out = slice + what is there.  A nonlinear epilogue runs on the LAST slice
        // only, with the earlier partial sums added inside it (kResidPre); a linear one takes the caller's resid on the first.
        const float* rs = first ? (resid ? resid + n0 : nullptr) : out + n0;
        const bool epi = nonlinear && last;
        if (rows_gemm_tc_launch(a + k0, ws, w_is_nk, (bias && first) ? bias + n0 : nullptr, epi ? relu : 0, (epi && gate) ? gate + n0 : nullptr,
                                rs, out + n0, R, 128, 128, (epi && !first) ? tc::kResidPre : 0, K, ldw, N, split ? 1 : 0, s))
          return 1;
      }
    return 0;
  }
  if ((flags & DG_OUT_BF16) && (resid || (gate && !(flags & DG_GATE_BF16))))
    return fail("dg_rows_gemm: a bf16 output takes no resid and only a bf16 gate");
  if (!(flags & DG_OUT_BF16) && (flags & DG_GATE_BF16)) return fail("dg_rows_gemm: a bf16 gate needs a bf16 output");
  return rows_gemm_tc_launch(a, w, w_is_nk, bias, relu, gate, resid, out, R, K, N, flags, K, w_is_nk ? K : N, N, 0, s);
}

static int gemm_tn_tc_launch(const float* a, const float* b, float* out, float* colsum_a, long long R, int M, int N, int flags,
                             int lda, int ldb, int ldo, int split, cudaStream_t s) {
  const int stage_bytes = (M + N) / 64 * tc::kTnBlk * (split ? 2 : 1);
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return fail("dg_gemm_tn: shape does not fit shared memory");
  const int smem = stages * stage_bytes + 4 * 32 * tc::kStage * 4 + 256 + 1024;
  {   // (per-device attribute, see rows_gemm_tc)
    cudaError_t e = cudaFuncSetAttribute(tc::gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute(gemm_tn_tc): %s", cudaGetErrorString(e));
  }
  long long tiles = (R + tc::kTnRows - 1) / tc::kTnRows;
  long long ctas = tiles < sm_count() ? tiles : sm_count();
  long long per = (tiles + ctas - 1) / ctas;
  // every CTA ends with an atomic flush of the whole [M,N] block: on node-sized inputs (R = B N rows, ~10 tiles per CTA) the
  // 148 flushes into the same 64-192 KB cost more than the contraction -- give a CTA at least 32 tiles (2048 rows)
  if (per < 32) per = tiles < 32 ? tiles : 32;
  ctas = (tiles + per - 1) / per;
  tc::gemm_tn_tc_kernel<<<(int)ctas, tc::kTnThreads, smem, s>>>(a, b, out, colsum_a, R, M, N, per, stages, flags,
                                                                opt_get(DG_OPT_L2_PREFETCH) & (M + N == 256 ? DG_PF_GEMM_TN : DG_PF_GEMM_TN_WIDE),
                                                                lda, ldb, ldo, split);
  return check_launch(split ? "dg_gemm_tn(bf16x3)" : "dg_gemm_tn(bf16)");
}

int gemm_tn_tc(const void* a_, const void* b_, float* out, float* colsum_a, long long R, int M, int N, int prec, int flags,
               cudaStream_t s) {
  const float* a = (const float*)a_;
  const float* b = (const float*)b_;
  const bool split = prec == DG_PREC_BF16X3;
  const bool blocks = M % 128 == 0 && N % 128 == 0 && M >= 128 && N >= 128;
  const bool ok = blocks && (M / 128) * N <= 512 && N <= 384 && M <= 384;
  const bool sliced = blocks && (split || (!ok && !flags));      // wider than one launch takes (H = 512): 128 x 128 output blocks
  if (!ok && !sliced) {
    if (flags) return fail("dg_gemm_tn: bf16 storage is only available for the tcgen05 shapes (M=%d N=%d)", M, N);
    return gemm_tn_fp32(a, b, out, colsum_a, R, M, N, s);
  }
  if ((!(flags & DG_A_BF16) && (reinterpret_cast<uintptr_t>(a) & 31)) || (!(flags & DG_OUT_BF16) && (reinterpret_cast<uintptr_t>(b) & 31)))
    return fail("dg_gemm_tn: fp32 operands must be 32-byte aligned (256-bit loads)");
  if (sliced) {     // 128 x 128 output blocks per launch (twice the operand bytes per stage when split); the bias gradient rides on the first N slice
    if (flags) return fail("dg_gemm_tn: the bf16x3 mode keeps every tensor fp32 (no bf16 storage flags)");
    for (int m0 = 0; m0 < M; m0 += 128)
      for (int n0 = 0; n0 < N; n0 += 128)
        if (gemm_tn_tc_launch(a + m0, b + n0, out + (long long)m0 * N + n0, (colsum_a && n0 == 0) ? colsum_a + m0 : nullptr, R, 128, 128, 0,
                              M, N, N, split ? 1 : 0, s))
          return 1;
    return 0;
  }
  return gemm_tn_tc_launch(a, b, out, colsum_a, R, M, N, flags, M, N, N, 0, s);
}

}  // namespace dg

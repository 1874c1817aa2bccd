// Fused residual-MLP forward of the encoder block (layers.py:41-54 + the residual and LayerNorm of
// layers.py:191/192) in ONE tcgen05 kernel:
//
//      out = LayerNorm( x + fc2( relu( fc1(x) + b1 ) ) + b2 ) * gamma + beta
//
// x:[R,128] fp32 rows (edge rows B*N*N, or node rows B*N), hidden width H = 128*HC (HC <= 3).
// The 128 x H hidden activation never leaves the SM: it goes TMEM -> registers (bias, ReLU, bf16)
// -> 128B-swizzled smem operand blocks -> second contraction.  HBM traffic is the algorithmic
// minimum: read x once (plus an L2-resident re-read for the fp32 residual), write out once.
//
// Persistent CTA per SM, 448 threads, warp-specialised:
//   warps 0-7   epilogue : thread = half a row (TMEM lane quarter w&3, column half w>>2);
//                          TMEM accumulator -> h chunk (bf16 operand block) / final residual + LayerNorm
//   warps 8-11  x loader : fp32 LDG.128 -> bf16 -> swizzled operand blocks (double-buffered tiles)
//   warp  12    MMA      : one thread issues tcgen05.mma; fc2 of tile t interleaved with fc1 of t+1
//   warp  13    W loader : one thread streams pre-packed bf16 weight stages (32 KB) with
//                          cp.async.bulk into a 2-stage ring (weights live in L2: 192 KB per net)
// TMEM: columns [0,384) three fc1 chunk accumulators, [384,512) the fc2 accumulator.
#include "tc_common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {
namespace tc {

constexpr int kMlpThreads = 448;         // warps 0-7 epilogue, 8-11 x loader, 12 MMA, 13 weight streamer
constexpr int kWStage = 2 * kBlkBytes;   // one packed weight stage: [2 kb][128 rows][128 B] = 32 KB
constexpr int kStgPitch = 20;            // epilogue transpose: 32 rows x 16 cols per warp, pitch 20 floats

// ---- weight pre-pack: fp32 nn.Linear weights -> bf16 swizzled operand stages in a workspace ------
// stage c        (c < HC): fc1 rows [c*128, c*128+128) of W1[H,128]        (B operand: N = hidden unit, K = in)
// stage HC + c           : fc2 columns [c*128, c*128+128) of W2[128,H]      (B operand: N = out, K = hidden slice)
__global__ void mlp_pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2, uint8_t* __restrict__ ws,
                                        int H) {
  const int HC = H / 128;
  const int total = 2 * HC * 128 * 16;          // (stage, row, kb*8 + j)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int c16 = idx & 15, row = (idx >> 4) & 127, stage = idx >> 11;
    int kb = c16 >> 3, j = c16 & 7;
    const float* src = stage < HC ? w1 + (long long)(stage * 128 + row) * 128 + kb * 64 + j * 8
                                  : w2 + (long long)row * H + (stage - HC) * 128 + kb * 64 + j * 8;
    st_block_chunk(ws + (long long)stage * kWStage + kb * kBlkBytes, row, j, ld4(src), ld4(src + 4));
  }
}

struct MlpSmem {
  static constexpr int xb = 0;                          // 2 x 32 KB
  static constexpr int hb = xb + 2 * kWStage;           // 2 x 32 KB
  static constexpr int wb = hb + 2 * kWStage;           // 2 x 32 KB
  static constexpr int stage = wb + 2 * kWStage;        // 8 warps x 32 x kStgPitch floats
  static constexpr int stats = stage + 8 * 32 * kStgPitch * 4;   // [2 parity][2 halves][128 rows] float2
  static constexpr int vec = stats + 2 * 2 * 128 * 8;   // b1[384] b2[128] gamma[128] beta[128]
  static constexpr int bars = vec + (384 + 3 * 128) * 4;
  static constexpr int total = bars + 256;
};

// warp-cooperative transposes through a [32][kStgPitch] staging tile (16 columns at a time):
// global rows -> "thread = row" registers, and back.  Global accesses are 64 B contiguous per row.
// gather = issue (all 16 coalesced LDG.128 of a 32-row x 64-col panel in flight at once) + finish (transpose)
__device__ __forceinline__ void gather_issue64(const float* __restrict__ src, long long row_base, long long R, int col,
                                               int lane, float4* xq) {
#pragma unroll
  for (int g16 = 0; g16 < 4; ++g16)
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2);
      xq[g16 * 4 + it] = (row_base + r < R) ? ld4(src + (row_base + r) * 128 + col + g16 * 16 + (lane & 3) * 4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__device__ __forceinline__ void gather_finish64(const float4* xq, float* stg, int lane, float* dst64) {
#pragma unroll
  for (int g16 = 0; g16 < 4; ++g16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) st4(stg + (it * 8 + (lane >> 2)) * kStgPitch + (lane & 3) * 4, xq[g16 * 4 + it]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = ld4(stg + lane * kStgPitch + i * 4);
      dst64[g16 * 16 + 4 * i] = v.x; dst64[g16 * 16 + 4 * i + 1] = v.y; dst64[g16 * 16 + 4 * i + 2] = v.z; dst64[g16 * 16 + 4 * i + 3] = v.w;
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void scatter_rows16(float* __restrict__ dst, long long row_base, long long R, int col, float* stg,
                                               int lane, const float* src16) {
#pragma unroll
  for (int i = 0; i < 4; ++i) st4(stg + lane * kStgPitch + i * 4, make_float4(src16[4 * i], src16[4 * i + 1], src16[4 * i + 2], src16[4 * i + 3]));
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    if (row_base + r < R) st4(dst + (row_base + r) * 128 + col + (lane & 3) * 4, ld4(stg + r * kStgPitch + (lane & 3) * 4));
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fwd_tc_kernel(const float* __restrict__ x, const uint8_t* __restrict__ wpack, const float* __restrict__ b1,
                  const float* __restrict__ b2, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float* __restrict__ out, long long R, int HC, float eps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem + MlpSmem::xb;
  uint8_t* sH = smem + MlpSmem::hb;
  uint8_t* sW = smem + MlpSmem::wb;
  float* sStage = reinterpret_cast<float*>(smem + MlpSmem::stage);
  float2* sStats = reinterpret_cast<float2*>(smem + MlpSmem::stats);
  float* sB1 = reinterpret_cast<float*>(smem + MlpSmem::vec);
  float* sB2 = sB1 + 384;
  float* sG = sB2 + 128;
  float* sBe = sG + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MlpSmem::bars);
  uint64_t *x_full = bars, *x_empty = bars + 2, *w_full = bars + 4, *w_empty = bars + 6, *hacc_full = bars + 8,
           *hacc_empty = bars + 11, *hb_full = bars + 14, *hb_empty = bars + 16, *z_full = bars + 18, *z_empty = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long num_tiles = (R + 127) / 128;
  const long long my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 128); mbar_init(&x_empty[i], 1);
      mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
      mbar_init(&hb_full[i], 256); mbar_init(&hb_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) { mbar_init(&hacc_full[i], 1); mbar_init(&hacc_empty[i], 256); }
    mbar_init(z_full, 1); mbar_init(z_empty, 256);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < HC * 128; i += kMlpThreads) sB1[i] = b1[i];
  for (int i = tid; i < 128; i += kMlpThreads) { sB2[i] = b2[i]; sG[i] = gamma[i]; sBe[i] = beta[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8 && warp < 12) {
    // ------------------------------------------------------------------ x loader
    const int lt = tid - 256;
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long row0 = (blockIdx.x + ti * gridDim.x) * 128;
      const int xs = ti & 1;
      mbar_wait(&x_empty[xs], ((ti >> 1) & 1) ^ 1);
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        uint8_t* blk = sX + xs * kWStage + kb * kBlkBytes;
        float4 v[16];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          int item = it * 128 + lt, r = item >> 3, j = item & 7;
          if (row0 + r < R) {
            const float* p = x + (row0 + r) * 128 + kb * 64 + j * 8;
            v[2 * it] = ld4(p); v[2 * it + 1] = ld4(p + 4);
          } else {
            v[2 * it] = v[2 * it + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          int item = it * 128 + lt;
          st_block_chunk(blk, item >> 3, item & 7, v[2 * it], v[2 * it + 1]);
        }
      }
      fence_async_smem();
      mbar_arrive(&x_full[xs]);
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ weight streamer (one thread)
    if (lane == 0 && my_tiles > 0) {
      uint32_t wcount = 0;
      auto push = [&](int stage) {
        const int ws = wcount & 1;
        mbar_wait(&w_empty[ws], ((wcount >> 1) & 1) ^ 1);
        mbar_expect_tx(&w_full[ws], kWStage);
        bulk_g2s(sW + ws * kWStage, wpack + (long long)stage * kWStage, kWStage, &w_full[ws]);
        ++wcount;
      };
      for (int c = 0; c < HC; ++c) push(c);                         // fc1 of the first tile
      for (long long ti = 0; ti < my_tiles; ++ti)
        for (int c = 0; c < HC; ++c) {
          push(HC + c);                                             // fc2 chunk c of tile ti
          if (ti + 1 < my_tiles) push(c);                           // fc1 chunk c of tile ti+1
        }
    }
    __syncwarp();
  } else if (warp == 12) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = make_idesc(128, 128, 0, 0);
      uint32_t wcount = 0, hcount = 0;
      auto mma_chunk = [&](uint32_t a_base, uint32_t d_col, bool first_clears) {
        const int ws = wcount & 1;
        mbar_wait(&w_full[ws], (wcount >> 1) & 1);
        tc_fence_after();
        const uint32_t b_base = smem_u32(sW + ws * kWStage);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * kBlkBytes + (kk & 3) * 32;
          umma_bf16(tmem_base + d_col, make_sdesc(a_base + off, 16, 1024), make_sdesc(b_base + off, 16, 1024), idesc,
                    (first_clears && kk == 0) ? 0u : 1u);
        }
        umma_commit(&w_empty[ws]);
        ++wcount;
      };
      auto fc1 = [&](long long ti, int c) {
        const int xs = ti & 1;
        if (c == 0) { mbar_wait(&x_full[xs], (ti >> 1) & 1); tc_fence_after(); }
        mbar_wait(&hacc_empty[c], (ti & 1) ^ 1);
        tc_fence_after();
        mma_chunk(smem_u32(sX + xs * kWStage), c * 128, true);
        umma_commit(&hacc_full[c]);
        if (c == HC - 1) umma_commit(&x_empty[xs]);
      };
      for (int c = 0; c < HC; ++c) fc1(0, c);
      for (long long ti = 0; ti < my_tiles; ++ti) {
        for (int c = 0; c < HC; ++c) {
          const int hs = hcount & 1;
          mbar_wait(&hb_full[hs], (hcount >> 1) & 1);
          if (c == 0) mbar_wait(z_empty, (ti & 1) ^ 1);
          tc_fence_after();
          mma_chunk(smem_u32(sH + hs * kWStage), 384, c == 0);
          umma_commit(&hb_empty[hs]);
          ++hcount;
          if (ti + 1 < my_tiles) fc1(ti + 1, c);
        }
        umma_commit(z_full);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-7)
    // warp w: TMEM lane quarter q = w & 3 (rows 32q..32q+31 of the tile), column half hf = w >> 2
    const int q = warp & 3, hf = warp >> 2;
    float* stg = sStage + warp * 32 * kStgPitch;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    uint32_t hcount = 0;
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long row0 = (blockIdx.x + ti * gridDim.x) * 128;
      for (int c = 0; c < HC; ++c) {
        mbar_wait(&hacc_full[c], ti & 1);
        const int hs = hcount & 1;
        mbar_wait(&hb_empty[hs], ((hcount >> 1) & 1) ^ 1);
        tc_fence_after();
        uint8_t* hblk = sH + hs * kWStage + hf * kBlkBytes;          // this half's 64 hidden columns = one operand block
        float v[64];
        tmem_ld32(tmem_base + lane_base + c * 128 + hf * 64, v);
        tmem_ld32(tmem_base + lane_base + c * 128 + hf * 64 + 32, v + 32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&hacc_empty[c]);
        const float* bb = sB1 + c * 128 + hf * 64;
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i] + bb[i], 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_block_chunk(hblk, row, j, make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
                         make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
        fence_async_smem();
        mbar_arrive(&hb_full[hs]);
        ++hcount;
      }
      // ---- final: a = z + b2 + x (this thread: one row, 64 columns) -> LayerNorm -> out
      float a[64];
      {
        float4 xq[16];
        gather_issue64(x, row0 + q * 32, R, hf * 64, lane, xq);
        gather_finish64(xq, stg, lane, a);
      }
      mbar_wait(z_full, ti & 1);
      tc_fence_after();
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
      for (int cgl = 0; cgl < 2; ++cgl) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + 384 + hf * 64 + cgl * 32, v);
        tmem_ld_wait();
        if (cgl == 1) { tc_fence_before(); mbar_arrive(z_empty); }
        const float* bb = sB2 + hf * 64 + cgl * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float t0 = a[cgl * 32 + i] + v[i] + bb[i], t1 = a[cgl * 32 + i + 1] + v[i + 1] + bb[i + 1];
          a[cgl * 32 + i] = t0; a[cgl * 32 + i + 1] = t1;
          s1a += t0; s1b += t1; s2a = fmaf(t0, t0, s2a); s2b = fmaf(t1, t1, s2b);
        }
      }
      float2* st = sStats + (ti & 1) * 256;
      st[hf * 128 + row] = make_float2(s1a + s1b, s2a + s2b);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 other = st[(hf ^ 1) * 128 + row];
      const float mean = (s1a + s1b + other.x) * (1.f / 128.f);
      const float rstd = rsqrtf(fmaxf((s2a + s2b + other.y) * (1.f / 128.f) - mean * mean, 0.f) + eps);
      const float* gg = sG + hf * 64;
      const float* be = sBe + hf * 64;
#pragma unroll
      for (int i = 0; i < 64; ++i) a[i] = (a[i] - mean) * rstd * gg[i] + be[i];
#pragma unroll
      for (int g16 = 0; g16 < 4; ++g16) scatter_rows16(out, row0 + q * 32, R, hf * 64 + g16 * 16, stg, lane, a + g16 * 16);
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
}  // namespace dg

using namespace dg;

extern "C" int dg_mlp_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* gamma, const float* beta, float* out, long long R, int D, int H, float eps,
                          void* workspace, long long workspace_bytes, void* stream) {
  if (R <= 0) return fail("dg_mlp_fwd: rows must be > 0");
  if (D != 128 || H % 128 || H < 128 || H > 384) return fail("dg_mlp_fwd: needs D == 128 and H in {128,256,384}, got D=%d H=%d", D, H);
  const int HC = H / 128;
  if (workspace_bytes < (long long)2 * HC * tc::kWStage) return fail("dg_mlp_fwd: workspace too small (%lld bytes)", workspace_bytes);
  if (reinterpret_cast<uintptr_t>(workspace) & 127) return fail("dg_mlp_fwd: workspace must be 128-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  tc::mlp_pack_weights_kernel<<<48, 256, 0, s>>>(w1, w2, (uint8_t*)workspace, H);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc::mlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::MlpSmem::total + 1024);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute(mlp_fwd): %s", cudaGetErrorString(e));
    configured = true;
  }
  long long tiles = (R + 127) / 128;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  tc::mlp_fwd_tc_kernel<<<grid, tc::kMlpThreads, tc::MlpSmem::total + 1024, s>>>(x, (const uint8_t*)workspace, b1, b2, gamma, beta,
                                                                               out, R, HC, eps);
  return check_launch("dg_mlp_fwd");
}

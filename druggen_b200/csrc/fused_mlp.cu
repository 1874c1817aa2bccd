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
// Persistent CTA per SM, 320 threads, warp-specialised:
//   warps 0-3  epilogue  : thread = one row; TMEM accumulator -> h chunk (bf16 operand) / final LN
//   warps 4-7  x loader  : fp32 LDG.128 -> bf16 -> swizzled operand blocks (double-buffered tiles)
//   warp  8    MMA       : one thread issues tcgen05.mma; fc2 of tile t interleaved with fc1 of t+1
//   warp  9    W loader  : one thread streams pre-packed bf16 weight stages (32 KB) with
//                          cp.async.bulk into a 2-stage ring (weights live in L2: 192 KB per net)
// TMEM: columns [0,384) three fc1 chunk accumulators, [384,512) the fc2 accumulator.
#include "tc_common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {
namespace tc {

constexpr int kMlpThreads = 320;
constexpr int kWStage = 2 * kBlkBytes;   // one packed weight stage: [2 kb][128 rows][128 B] = 32 KB

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
  static constexpr int stage = wb + 2 * kWStage;        // 4 warps x 32 x 36 floats
  static constexpr int vec = stage + 4 * 32 * 36 * 4;   // b1[384] b2[128] gamma[128] beta[128]
  static constexpr int bars = vec + (384 + 3 * 128) * 4;
  static constexpr int total = bars + 256;
};

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
      mbar_init(&hb_full[i], 128); mbar_init(&hb_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) { mbar_init(&hacc_full[i], 1); mbar_init(&hacc_empty[i], 128); }
    mbar_init(z_full, 1); mbar_init(z_empty, 128);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < HC * 128; i += kMlpThreads) sB1[i] = b1[i];
  for (int i = tid; i < 128; i += kMlpThreads) { sB2[i] = b2[i]; sG[i] = gamma[i]; sBe[i] = beta[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ x loader
    const int lt = tid - 128;
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
  } else if (warp == 9) {
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
  } else if (warp == 8) {
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
    // ------------------------------------------------------------------ epilogue (warps 0-3), thread = row
    float* stg = sStage + warp * 32 * 36;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int row = warp * 32 + lane;
    uint32_t hcount = 0;
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long row0 = (blockIdx.x + ti * gridDim.x) * 128;
      for (int c = 0; c < HC; ++c) {
        mbar_wait(&hacc_full[c], ti & 1);
        const int hs = hcount & 1;
        mbar_wait(&hb_empty[hs], ((hcount >> 1) & 1) ^ 1);
        tc_fence_after();
        uint8_t* hblk = sH + hs * kWStage;
#pragma unroll 1
        for (int cg = 0; cg < 4; ++cg) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + c * 128 + cg * 32, v);
          tmem_ld_wait();
          if (cg == 3) { tc_fence_before(); mbar_arrive(&hacc_empty[c]); }
          const float* bb = sB1 + c * 128 + cg * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bb[i], 0.f);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_block_chunk(hblk + (cg >> 1) * kBlkBytes, row, (cg & 1) * 4 + q, make_float4(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]),
                           make_float4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]));
        }
        fence_async_smem();
        mbar_arrive(&hb_full[hs]);
        ++hcount;
      }
      // ---- final: z + b2 + x -> LayerNorm -> out
      mbar_wait(z_full, ti & 1);
      tc_fence_after();
      const long long grow = row0 + row;
      const bool live = grow < R;
      const float* xr = x + (live ? grow : 0) * 128;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int cg = 0; cg < 4; ++cg) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + 384 + cg * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 xv = ld4(xr + cg * 32 + i * 4);
          float a0 = v[4 * i] + sB2[cg * 32 + 4 * i] + xv.x, a1 = v[4 * i + 1] + sB2[cg * 32 + 4 * i + 1] + xv.y;
          float a2 = v[4 * i + 2] + sB2[cg * 32 + 4 * i + 2] + xv.z, a3 = v[4 * i + 3] + sB2[cg * 32 + 4 * i + 3] + xv.w;
          s1 += (a0 + a1) + (a2 + a3);
          s2 += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        }
      }
      const float mean = s1 * (1.f / 128.f);
      const float rstd = rsqrtf(fmaxf(s2 * (1.f / 128.f) - mean * mean, 0.f) + eps);
#pragma unroll 1
      for (int cg = 0; cg < 4; ++cg) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + 384 + cg * 32, v);
        tmem_ld_wait();
        if (cg == 3) { tc_fence_before(); mbar_arrive(z_empty); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 xv = ld4(xr + cg * 32 + i * 4);
          const int cc = cg * 32 + 4 * i;
          float4 o;
          o.x = (v[4 * i] + sB2[cc] + xv.x - mean) * rstd * sG[cc] + sBe[cc];
          o.y = (v[4 * i + 1] + sB2[cc + 1] + xv.y - mean) * rstd * sG[cc + 1] + sBe[cc + 1];
          o.z = (v[4 * i + 2] + sB2[cc + 2] + xv.z - mean) * rstd * sG[cc + 2] + sBe[cc + 2];
          o.w = (v[4 * i + 3] + sB2[cc + 3] + xv.w - mean) * rstd * sG[cc + 3] + sBe[cc + 3];
          st4(stg + lane * 36 + i * 4, o);
        }
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = rr * 4 + (lane >> 3);
          const long long gr = row0 + warp * 32 + r;
          if (gr < R) st4(out + gr * 128 + cg * 32 + (lane & 7) * 4, ld4(stg + r * 36 + (lane & 7) * 4));
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
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

// fp32 CUDA-core contractions (DG_PREC_FP32, the parity mode): plain fp32 FMA with fp32
// accumulation, shared-memory tiled, register-blocked.  They exist so that every result of the
// tensor-core paths can be checked on the device against fp32 arithmetic at full problem size,
// and so the 1e-3 parity bar is met with ~1e-6 to spare; the throughput mode is gemm_tc.cu.
//
//   rows_gemm : out[R,N] = epi(a[R,K] . op(w) + bias)      128x64 CTA tile, 8x4 per thread
//   gemm_tn   : out[M,N] += a[R,M]^T . b[R,N]              64x64 tile, split over rows, atomics
#include "common.cuh"
#include "../../include/druggen_b200.h"

namespace dg {

constexpr int BM = 128, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
rows_gemm_fp32_kernel(const float* __restrict__ a, const float* __restrict__ w, int w_is_nk,
                      const float* __restrict__ bias, int relu, const float* __restrict__ gate,
                      const float* __restrict__ resid, float* __restrict__ out, long long R, int K, int N) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;               // 16 x 16 threads; thread tile 8 rows x 4 cols
  const long long row0 = (long long)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: 128 rows x 16 k  = 512 float4, two per thread, stored transposed
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int f = tid + it * 256;
      int r = f >> 2, kq = (f & 3) * 4;
      long long gr = row0 + r;
      float4 v = make_float4(0, 0, 0, 0);
      if (gr < R && k0 + kq < K) v = ld4(a + gr * K + k0 + kq);
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
    if (w_is_nk) {              // w[N][K]: 64 n x 16 k = 256 float4 along k
      int n = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0, 0, 0, 0);
      if (col0 + n < N && k0 + kq < K) v = ld4(w + (long long)(col0 + n) * K + k0 + kq);
      Ws[kq + 0][n] = v.x; Ws[kq + 1][n] = v.y; Ws[kq + 2][n] = v.z; Ws[kq + 3][n] = v.w;
    } else {                    // w[K][N]: 16 k x 64 n = 256 float4 along n
      int kk = tid >> 4, nq = (tid & 15) * 4;
      float4 v = make_float4(0, 0, 0, 0);
      if (k0 + kk < K && col0 + nq < N) v = ld4(w + (long long)(k0 + kk) * N + col0 + nq);
      *reinterpret_cast<float4*>(&Ws[kk][nq]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int c = col0 + tx * 4;
  if (c >= N) return;
  float4 bz = bias ? ld4(bias + c) : make_float4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long r = row0 + ty * 8 + i;
    if (r >= R) break;
    float4 o = make_float4(acc[i][0] + bz.x, acc[i][1] + bz.y, acc[i][2] + bz.z, acc[i][3] + bz.w);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (gate) {
      float4 g = ld4(gate + r * N + c);
      o.x = g.x > 0.f ? o.x : 0.f; o.y = g.y > 0.f ? o.y : 0.f; o.z = g.z > 0.f ? o.z : 0.f; o.w = g.w > 0.f ? o.w : 0.f;
    }
    if (resid) {
      float4 z = ld4(resid + r * N + c);
      o.x += z.x; o.y += z.y; o.z += z.z; o.w += z.w;
    }
    st4(out + r * N + c, o);
  }
}

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
gemm_tn_fp32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                    float* __restrict__ colsum_a, long long R, int M, int N, long long rows_per_split) {
  __shared__ float As[TK][TM];
  __shared__ float Bs[TK][TN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  long long r0 = (long long)blockIdx.z * rows_per_split;
  long long r1 = r0 + rows_per_split < R ? r0 + rows_per_split : R;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 4, lc = (tid & 15) * 4;          // 16 rows x 16 float4
  const bool do_colsum = colsum_a != nullptr && blockIdx.y == 0;   // one N-tile column of CTAs sums a's columns
  float csum = 0.f;
  for (long long r = r0; r < r1; r += TK) {
    float4 va = make_float4(0, 0, 0, 0), vb = make_float4(0, 0, 0, 0);
    if (r + lr < r1) {
      if (m0 + lc < M) va = ld4(a + (r + lr) * M + m0 + lc);
      if (n0 + lc < N) vb = ld4(b + (r + lr) * N + n0 + lc);
    }
    *reinterpret_cast<float4*>(&As[lr][lc]) = va;
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = vb;
    __syncthreads();
    if (do_colsum && tid < TM) {
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) csum += As[kk][tid];
    }
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float4 x = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 y = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float xv[4] = {x.x, x.y, x.z, x.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (do_colsum && tid < TM && m0 + tid < M) atomicAdd(colsum_a + m0 + tid, csum);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) atomicAdd(out + (long long)m * N + n, acc[i][j]);
    }
  }
}

int rows_gemm_fp32(const float* a, const float* w, int w_is_nk, const float* bias, int relu, const float* gate,
                   const float* resid, float* out, long long R, int K, int N, cudaStream_t s) {
  if (K % 4 || N % 4) return fail("rows_gemm(fp32): K=%d and N=%d must be multiples of 4", K, N);
  long long gx = (R + BM - 1) / BM;
  if (gx > 2147483647LL) return fail("too many rows");
  dim3 grid((unsigned)gx, (N + BN - 1) / BN);
  rows_gemm_fp32_kernel<<<grid, 256, 0, s>>>(a, w, w_is_nk, bias, relu, gate, resid, out, R, K, N);
  return check_launch("dg_rows_gemm(fp32)");
}

int gemm_tn_fp32(const float* a, const float* b, float* out, float* colsum_a, long long R, int M, int N, cudaStream_t s) {
  if (M % 4 || N % 4) return fail("gemm_tn(fp32): M=%d and N=%d must be multiples of 4", M, N);
  int tiles = ((M + TM - 1) / TM) * ((N + TN - 1) / TN);
  long long splits = (sm_count() * 4 + tiles - 1) / tiles;
  long long max_splits = (R + 4 * TK - 1) / (4 * TK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long rps = (R + splits - 1) / splits;
  rps = (rps + TK - 1) / TK * TK;
  splits = (R + rps - 1) / rps;
  dim3 grid((M + TM - 1) / TM, (N + TN - 1) / TN, (unsigned)splits);
  gemm_tn_fp32_kernel<<<grid, 256, 0, s>>>(a, b, out, colsum_a, R, M, N, rps);
  return check_launch("dg_gemm_tn(fp32)");
}

}  // namespace dg

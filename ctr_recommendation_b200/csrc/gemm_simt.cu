// fp32 FMA GEMM (FBN_PREC_FP32): the exact-arithmetic mode of the dense contractions and the
// validation twin of the tcgen05 path in gemm_tc.cu.  128x128x16 tiles, 256 threads, 8x8 register
// micro-tiles, register-prefetch double buffering, optional split-K / batching through blockIdx.z,
// 128-wide K-block and N-block skip masks (the MLP input has structurally zero blocks because the
// user field is the constant zero vector, src/model_fibinet.py:152).
#include "common.cuh"
#include "gemm.h"

namespace fbn {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool A_T, bool B_T>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int z = blockIdx.z;
  const int bi = z / g.splits, sp = z % g.splits;
  const long long m0 = (long long)blockIdx.y * BM;
  const long long n0 = (long long)blockIdx.x * BN;
  if (g.nmask != ~0ull && !((g.nmask >> (n0 / 128)) & 1ull)) return;
  const float* A = g.A + bi * g.strideA;
  const float* Bm = g.B + bi * g.strideB;
  float* C = g.C + bi * g.strideC + sp * g.strideSplit;
  const long long ktiles = (g.K + BK - 1) / BK;
  const long long per = (ktiles + g.splits - 1) / g.splits;
  const long long kt0 = sp * per, kt1 = min(ktiles, kt0 + per);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto active = [&](long long kt) { return g.kmask == ~0ull || ((g.kmask >> ((kt * BK) / 128)) & 1ull); };
  auto gload = [&](long long kt) {
    const long long k0 = kt * BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      if (A_T) {  // stored (K,M): float4 along m
        const int k = idx >> 5, mq = idx & 31;
        const long long kk = k0 + k, mm = m0 + mq * 4;
        ra[i] = (kk < g.K && mm < g.M) ? ld4(A + kk * g.lda + mm) : f4(0.f);
      } else {    // stored (M,K): float4 along k
        const int r = idx >> 2, kq = idx & 3;
        const long long mm = m0 + r, kk = k0 + kq * 4;
        ra[i] = (mm < g.M && kk < g.K) ? ld4(A + mm * g.lda + kk) : f4(0.f);
      }
      if (!B_T) {  // stored (K,N): float4 along n
        const int k = idx >> 5, nq = idx & 31;
        const long long kk = k0 + k, nn = n0 + nq * 4;
        rb[i] = (kk < g.K && nn < g.N) ? ld4(Bm + kk * g.ldb + nn) : f4(0.f);
      } else {     // stored (N,K): float4 along k
        const int r = idx >> 2, kq = idx & 3;
        const long long nn = n0 + r, kk = k0 + kq * 4;
        rb[i] = (nn < g.N && kk < g.K) ? ld4(Bm + nn * g.ldb + kk) : f4(0.f);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      if (A_T) {
        const int k = idx >> 5, mq = idx & 31;
        *reinterpret_cast<float4*>(&As[buf][k][mq * 4]) = ra[i];
      } else {
        const int r = idx >> 2, kq = idx & 3;
        As[buf][kq * 4 + 0][r] = ra[i].x; As[buf][kq * 4 + 1][r] = ra[i].y;
        As[buf][kq * 4 + 2][r] = ra[i].z; As[buf][kq * 4 + 3][r] = ra[i].w;
      }
      if (!B_T) {
        const int k = idx >> 5, nq = idx & 31;
        *reinterpret_cast<float4*>(&Bs[buf][k][nq * 4]) = rb[i];
      } else {
        const int r = idx >> 2, kq = idx & 3;
        Bs[buf][kq * 4 + 0][r] = rb[i].x; Bs[buf][kq * 4 + 1][r] = rb[i].y;
        Bs[buf][kq * 4 + 2][r] = rb[i].z; Bs[buf][kq * 4 + 3][r] = rb[i].w;
      }
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  long long kt = kt0;
  while (kt < kt1 && !active(kt)) ++kt;
  int buf = 0;
  if (kt < kt1) { gload(kt); sstore(0); }
  __syncthreads();
  while (kt < kt1) {
    long long nxt = kt + 1;
    while (nxt < kt1 && !active(nxt)) ++nxt;
    if (nxt < kt1) gload(nxt);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (nxt < kt1) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
    kt = nxt;
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const long long n = n0 + jh * 64 + tx * 4;
      if (n >= g.N) continue;
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (g.bias) v += ld4(g.bias + n);
      float* cp = C + m * g.ldc + n;
      if (g.accumulate) v += *reinterpret_cast<const float4*>(cp);
      st4(cp, v);
    }
  }
}

int gemm_simt(const GemmArgs& g, cudaStream_t st) {
  FBN_REQUIRE(g.N % 4 == 0 && g.lda % 4 == 0 && g.ldb % 4 == 0 && g.ldc % 4 == 0, FBN_ERR_SHAPE,
              "gemm: N and leading dimensions must be multiples of 4");
  FBN_REQUIRE((g.a_t || g.K % 4 == 0) && (!g.b_t || g.K % 4 == 0), FBN_ERR_SHAPE, "gemm: K must be a multiple of 4");
  FBN_REQUIRE(!g.a_t || g.M % 4 == 0, FBN_ERR_SHAPE, "gemm: M must be a multiple of 4 for transposed A");
  if (g.M <= 0 || g.N <= 0) return FBN_OK;
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), (unsigned)(g.batch * g.splits));
  if (g.a_t && g.b_t) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else if (g.a_t) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (g.b_t) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

}  // namespace fbn

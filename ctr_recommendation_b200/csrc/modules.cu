// Stand-alone SENetLayer / BilinearInteraction entry points (module-level API of the reference,
// src/model_fibinet.py:5-89) for arbitrary field counts (F <= 64, hidden <= 64, D % 4 == 0).  The *_ld variants take row
// strides so that the F-field model (general.py) can keep V and the pair products inside one (B, (F+P)*D) MLP-input buffer.
#include <algorithm>

#include "common.cuh"
#include "gemm.h"
#include "tower.h"

namespace fbn {

constexpr int SE_MAXF = 64;

// ---- SENET forward: warp per sample -----------------------------------------------------------
__global__ void __launch_bounds__(256) senet_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, long long B, int F, int Dm, int R,
                                                        float* __restrict__ y, long long ldy, float* __restrict__ gate) {
  __shared__ float zs[8][SE_MAXF], hs[8][SE_MAXF], ss[8][SE_MAXF];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); b < B; b += nw) {
    const float* xb = x + b * F * Dm;
    for (int f = 0; f < F; ++f) {                       // squeeze: mean over the embedding dim (ref :28)
      float t = 0.f;
      for (int d = lane * 4; d < Dm; d += 128) t += hsum4(ld4(xb + f * Dm + d));
      t = warp_sum(t);
      if (lane == 0) zs[w][f] = t / (float)Dm;
    }
    __syncwarp();
    for (int r = lane; r < R; r += 32) {                // excitation (ref :17-21)
      float a = b1[r];
      for (int f = 0; f < F; ++f) a = fmaf(zs[w][f], w1[r * F + f], a);
      hs[w][r] = fmaxf(a, 0.f);
    }
    __syncwarp();
    for (int f = lane; f < F; f += 32) {
      float a = b2[f];
      for (int r = 0; r < R; ++r) a = fmaf(hs[w][r], w2[f * R + r], a);
      const float s = sigmoidf_(a);
      ss[w][f] = s;
      gate[b * F + f] = s;
    }
    __syncwarp();
    for (int f = 0; f < F; ++f) {                       // re-weight (ref :35)
      const float s = ss[w][f];
      for (int d = lane * 4; d < Dm; d += 128) st4(y + b * ldy + f * Dm + d, ld4(xb + f * Dm + d) * s);
    }
    __syncwarp();
  }
}

// ---- SENET backward: dx + per-sample records {da2[F], h[R], z[F], da1[R]} -----------------------
__global__ void __launch_bounds__(256) senet_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ dy, long long B, int F,
                                                        int Dm, int R, float* __restrict__ dx, float* __restrict__ rec) {
  __shared__ float zs[8][SE_MAXF], hs[8][SE_MAXF], da2[8][SE_MAXF], da1[8][SE_MAXF], dz[8][SE_MAXF];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int RS = 2 * F + 2 * R;
  for (long long b = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); b < B; b += nw) {
    const float* xb = x + b * F * Dm;
    const float* db = dy + b * F * Dm;
    for (int f = 0; f < F; ++f) {
      float t = 0.f, u = 0.f;
      for (int d = lane * 4; d < Dm; d += 128) {
        const float4 xv = ld4(xb + f * Dm + d);
        t += hsum4(xv);
        u += hsum4(xv * ld4(db + f * Dm + d));
      }
      t = warp_sum(t);
      u = warp_sum(u);
      if (lane == 0) {
        const float s = gate[b * F + f];
        zs[w][f] = t / (float)Dm;
        da2[w][f] = u * s * (1.0f - s);
      }
    }
    __syncwarp();
    for (int r = lane; r < R; r += 32) {
      float a = b1[r];
      for (int f = 0; f < F; ++f) a = fmaf(zs[w][f], w1[r * F + f], a);
      const float h = fmaxf(a, 0.f);
      float t = 0.f;
      for (int f = 0; f < F; ++f) t = fmaf(da2[w][f], w2[f * R + r], t);
      hs[w][r] = h;
      da1[w][r] = h > 0.f ? t : 0.f;
    }
    __syncwarp();
    for (int f = lane; f < F; f += 32) {
      float t = 0.f;
      for (int r = 0; r < R; ++r) t = fmaf(da1[w][r], w1[r * F + f], t);
      dz[w][f] = t / (float)Dm;
    }
    __syncwarp();
    float* rb = rec + b * RS;
    for (int f = lane; f < F; f += 32) { rb[f] = da2[w][f]; rb[F + R + f] = zs[w][f]; }
    for (int r = lane; r < R; r += 32) { rb[F + r] = hs[w][r]; rb[2 * F + R + r] = da1[w][r]; }
    for (int f = 0; f < F; ++f) {
      const float s = gate[b * F + f], z = dz[w][f];
      for (int d = lane * 4; d < Dm; d += 128) st4(dx + b * F * Dm + f * Dm + d, ld4(db + f * Dm + d) * s + f4(z));
    }
    __syncwarp();
  }
}

// outputs t in [0, 2FR+F+R): dW1[r][f] (R*F) | db1 (R) | dW2[f][r] (F*R) | db2 (F)
__global__ void senet_pgrad_partial_kernel(const float* __restrict__ rec, long long B, long long per, int F, int R,
                                           float* __restrict__ partial) {
  const int nout = 2 * F * R + F + R, RS = 2 * F + 2 * R;
  const long long b0 = (long long)blockIdx.x * per, b1 = min(B, b0 + per);
  for (int t = threadIdx.x; t < nout; t += blockDim.x) {
    int ia, ib;
    if (t < R * F) { ia = 2 * F + R + t / F; ib = F + R + t % F; }
    else if (t < R * F + R) { ia = 2 * F + R + (t - R * F); ib = -1; }
    else if (t < 2 * R * F + R) { const int u = t - R * F - R; ia = u / R; ib = F + u % R; }
    else { ia = t - 2 * R * F - R; ib = -1; }
    float acc = 0.f;
    for (long long b = b0; b < b1; ++b) {
      const float* r = rec + b * RS;
      acc += ib >= 0 ? r[ia] * r[ib] : r[ia];
    }
    partial[(long long)blockIdx.x * nout + t] = acc;
  }
}

__global__ void __launch_bounds__(256) senet_pgrad_final_kernel(const float* __restrict__ partial, int parts, int F, int R, float* dw1,
                                                                float* db1, float* dw2, float* db2) {
  const int nout = 2 * F * R + F + R;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= nout) return;
  double acc = 0.0;
  for (int p = lane; p < parts; p += 32) acc += (double)partial[(long long)p * nout + t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane != 0) return;
  if (t < R * F) dw1[t] = (float)acc;
  else if (t < R * F + R) db1[t - R * F] = (float)acc;
  else if (t < 2 * R * F + R) dw2[t - R * F - R] = (float)acc;
  else db2[t - 2 * R * F - R] = (float)acc;
}

// ---- bilinear pair products for arbitrary F -----------------------------------------------------
__host__ __device__ inline int pair_index(int i, int j, int F) { return i * (2 * F - i - 1) / 2 + (j - i - 1); }

// forward: p[b][q(i,j)] = left * right with (ALL) v_i * T_j, (EACH) T_i * v_j, (INTERACTION) T_q * v_j
__global__ void bil_pairs_fwd_kernel(const float* __restrict__ v, long long ldv, const float* __restrict__ T, int type, long long B,
                                     int F, int Dm, float* __restrict__ p, long long ldp) {
  const int P = F * (F - 1) / 2, nT = type == FBN_BILINEAR_ALL ? F : (type == FBN_BILINEAR_EACH ? F - 1 : P);
  const int d4 = Dm / 4;
  const long long total = B * P * d4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % d4) * 4;
    const long long bq = idx / d4;
    const int q = (int)(bq % P);
    const long long b = bq / P;
    int i = 0, rem = q;
    while (rem >= F - 1 - i) { rem -= F - 1 - i; ++i; }
    const int j = i + 1 + rem;
    const float* vb = v + b * ldv;
    const float* Tb = T + b * nT * Dm;
    float4 o;
    if (type == FBN_BILINEAR_ALL) o = ld4(vb + i * Dm + d) * ld4(Tb + j * Dm + d);
    else if (type == FBN_BILINEAR_EACH) o = ld4(Tb + i * Dm + d) * ld4(vb + j * Dm + d);
    else o = ld4(Tb + q * Dm + d) * ld4(vb + j * Dm + d);
    st4(p + b * ldp + (long long)q * Dm + d, o);
  }
}

// backward (elementwise part): dT and the direct dv terms; thread per (b, field f, d4)
__global__ void bil_pairs_bwd_kernel(const float* __restrict__ v, long long ldv, const float* __restrict__ T, const float* __restrict__ dp,
                                     long long lddp, const float* __restrict__ dv_init, long long ldi, int type, long long B, int F,
                                     int Dm, float* __restrict__ dv, float* __restrict__ dT) {
  const int P = F * (F - 1) / 2, nT = type == FBN_BILINEAR_ALL ? F : (type == FBN_BILINEAR_EACH ? F - 1 : P);
  const int d4 = Dm / 4;
  const long long total = B * F * d4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % d4) * 4;
    const long long bf = idx / d4;
    const int f = (int)(bf % F);
    const long long b = bf / F;
    const float* vb = v + b * ldv;
    const float* Tb = T + b * nT * Dm;
    const float* dpb = dp + b * lddp;
    float4 gv = dv_init ? ld4(dv_init + b * ldi + f * Dm + d) : f4(0.f), gt = f4(0.f);   // dv_init: gradient that reached v directly
    if (type == FBN_BILINEAR_ALL) {
      for (int j = f + 1; j < F; ++j) gv += ld4(dpb + pair_index(f, j, F) * Dm + d) * ld4(Tb + j * Dm + d);   // as v_i
      for (int i = 0; i < f; ++i) gt += ld4(dpb + pair_index(i, f, F) * Dm + d) * ld4(vb + i * Dm + d);       // as T_j
      st4(dT + (b * nT + f) * Dm + d, gt);
    } else if (type == FBN_BILINEAR_EACH) {
      for (int i = 0; i < f; ++i) gv += ld4(dpb + pair_index(i, f, F) * Dm + d) * ld4(Tb + i * Dm + d);       // as v_j
      if (f < F - 1) {
        for (int j = f + 1; j < F; ++j) gt += ld4(dpb + pair_index(f, j, F) * Dm + d) * ld4(vb + j * Dm + d); // as T_i
        st4(dT + (b * nT + f) * Dm + d, gt);
      }
    } else {
      for (int i = 0; i < f; ++i) {
        const int q = pair_index(i, f, F);
        gv += ld4(dpb + q * Dm + d) * ld4(Tb + q * Dm + d);                                                 // as v_j
      }
      for (int j = f + 1; j < F; ++j) {
        const int q = pair_index(f, j, F);
        st4(dT + (b * nT + q) * Dm + d, ld4(dpb + q * Dm + d) * ld4(vb + j * Dm + d));                        // dT_q
      }
    }
    st4(dv + (b * F + f) * Dm + d, gv);
  }
}

static int ew_blocks(long long n) { return (int)std::max<long long>(1, std::min<long long>(cdiv(n, 256), 8LL * num_sms())); }

}  // namespace fbn

using namespace fbn;

extern "C" size_t fbn_senet_scratch_bytes(int64_t batch, int fields, int hidden) {
  const size_t nout = (size_t)2 * fields * hidden + fields + hidden;
  return ((size_t)batch * (2 * fields + 2 * hidden) + (size_t)1024 * nout) * sizeof(float) + 512;
}

extern "C" int fbn_senet_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t batch, int fields,
                             int dim, int hidden, float* y, float* gate, fbn_stream_t stream) {
  FBN_REQUIRE(x && w1 && b1 && w2 && b2 && y && gate, FBN_ERR_ARG, "fbn_senet_fwd: null pointer");
  FBN_REQUIRE(fields >= 1 && fields <= SE_MAXF && hidden >= 1 && hidden <= SE_MAXF && dim % 4 == 0 && dim >= 4, FBN_ERR_SHAPE,
              "fbn_senet_fwd: need 1 <= fields, hidden <= 64 and dim %% 4 == 0");
  FBN_REQUIRE(aligned16(x) && aligned16(y), FBN_ERR_ALIGN, "fbn_senet_fwd: unaligned pointer");
  if (batch <= 0) return FBN_OK;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(batch, 8), 8LL * num_sms()));
  senet_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, w1, b1, w2, b2, batch, fields, dim, hidden, y, (long long)fields * dim, gate);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_senet_bwd(const float* x, const float* gate, const float* w1, const float* b1, const float* w2, const float* dy,
                             int64_t batch, int fields, int dim, int hidden, float* dx, float* dw1, float* db1, float* dw2, float* db2,
                             void* scratch, size_t scratch_bytes, fbn_stream_t stream) {
  FBN_REQUIRE(x && gate && w1 && b1 && w2 && dy && dx && dw1 && db1 && dw2 && db2 && scratch, FBN_ERR_ARG, "fbn_senet_bwd: null pointer");
  FBN_REQUIRE(fields >= 1 && fields <= SE_MAXF && hidden >= 1 && hidden <= SE_MAXF && dim % 4 == 0, FBN_ERR_SHAPE, "fbn_senet_bwd: bad shape");
  FBN_REQUIRE(scratch_bytes >= fbn_senet_scratch_bytes(batch, fields, hidden), FBN_ERR_ARG, "fbn_senet_bwd: scratch too small");
  FBN_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx), FBN_ERR_ALIGN, "fbn_senet_bwd: unaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int nout = 2 * fields * hidden + fields + hidden;
  float* rec = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
  float* partial = rec + (size_t)batch * (2 * fields + 2 * hidden);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(batch, 8), 8LL * num_sms()));
  senet_bwd_kernel<<<blocks, 256, 0, st>>>(x, gate, w1, b1, w2, dy, batch, fields, dim, hidden, dx, rec);
  FBN_CHECK_LAUNCH();
  int parts = (int)std::max<long long>(1, std::min<long long>(cdiv(batch, 32), 1024));
  const long long per = cdiv(batch, parts);
  senet_pgrad_partial_kernel<<<parts, 128, 0, st>>>(rec, batch, per, fields, hidden, partial);
  FBN_CHECK_LAUNCH();
  senet_pgrad_final_kernel<<<(nout + 7) / 8, 256, 0, st>>>(partial, parts, fields, hidden, dw1, db1, dw2, db2);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// scratch layout: T (B,nT,D) | dT (B,nT,D) | split-K partials (nW * 32 * D * D) | tcgen05 operand scratch
static void bil_sizes(int type, int F, int& nT, int& nW) {
  const int P = F * (F - 1) / 2;
  nT = type == FBN_BILINEAR_ALL ? F : (type == FBN_BILINEAR_EACH ? F - 1 : P);
  nW = type == FBN_BILINEAR_ALL ? 1 : nT;
}

extern "C" size_t fbn_bilinear_scratch_bytes(int64_t batch, int fields, int dim, int type) {
  int nT, nW;
  bil_sizes(type, fields, nT, nW);
  const size_t tc = std::max(gemm_tc_scratch_bytes(batch * fields, dim, dim, FBN_PREC_TF32X3),
                             gemm_tc_scratch_bytes(dim, dim, batch * fields, FBN_PREC_TF32X3));
  return ((size_t)2 * batch * nT * dim + (size_t)32 * std::max(nW, 1) * dim * dim) * sizeof(float) + tc + 4096;
}

struct BilScratch { float* T; float* dT; float* partial; void* tc; size_t tc_bytes; };

static BilScratch bil_carve(void* scratch, size_t bytes, int64_t batch, int fields, int dim, int type) {
  int nT, nW;
  bil_sizes(type, fields, nT, nW);
  BilScratch s;
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~uintptr_t(1023));
  auto up = [](size_t x) { return (x + 1023) & ~size_t(1023); };
  s.T = (float*)p; p += up((size_t)batch * nT * dim * 4);
  s.dT = (float*)p; p += up((size_t)batch * nT * dim * 4);
  s.partial = (float*)p; p += up((size_t)32 * std::max(nW, 1) * dim * dim * 4);
  s.tc = p;
  s.tc_bytes = (size_t)((char*)scratch + bytes - p);
  return s;
}

// T = v_src * W_idx
static int bil_transform(const float* v, const float* w, int type, int64_t B, int F, int Dm, float* T, int precision, BilScratch& s,
                         cudaStream_t st) {
  GemmArgs g;
  g.N = Dm; g.K = Dm; g.ldb = Dm; g.b_t = 0;
  if (type == FBN_BILINEAR_ALL) {            // vid = v.view(B*F, D) @ W   (ref :72)
    g.A = v; g.lda = Dm; g.B = w; g.C = T; g.ldc = Dm; g.M = B * F;
    return gemm(g, precision, s.tc, s.tc_bytes, st);
  }
  g.M = B; g.lda = (long long)F * Dm;
  if (type == FBN_BILINEAR_EACH) {           // T_i = v_i W_i   (ref :85)
    g.A = v; g.strideA = Dm; g.B = w; g.strideB = (long long)Dm * Dm; g.C = T; g.ldc = (long long)(F - 1) * Dm; g.strideC = Dm; g.batch = F - 1;
    return gemm(g, precision, s.tc, s.tc_bytes, st);
  }
  const int P = F * (F - 1) / 2;
  for (int i = 0; i < F - 1; ++i) {          // T_q = v_i W_q
    const int q0 = pair_index(i, i + 1, F), nj = F - 1 - i;
    g.A = v + i * Dm; g.strideA = 0; g.B = w + (long long)q0 * Dm * Dm; g.strideB = (long long)Dm * Dm; g.C = T + q0 * Dm;
    g.ldc = (long long)P * Dm; g.strideC = Dm; g.batch = nj;
    int rc = gemm(g, precision, s.tc, s.tc_bytes, st);
    if (rc) return rc;
  }
  return FBN_OK;
}

static int bil_check(const float* v, const float* w, int type, int64_t batch, int fields, int dim, int precision) {
  FBN_REQUIRE(v && w, FBN_ERR_ARG, "bilinear: null pointer");
  FBN_REQUIRE(type >= FBN_BILINEAR_ALL && type <= FBN_BILINEAR_INTERACTION, FBN_ERR_ARG, "bilinear_type must be 'all' or 'each'");
  FBN_REQUIRE(fields >= 2 && fields <= 64 && dim % 4 == 0 && dim >= 4 && batch >= 1, FBN_ERR_SHAPE, "bilinear: bad shape");
  FBN_REQUIRE(precision == FBN_PREC_FP32 || dim % 128 == 0, FBN_ERR_SHAPE, "bilinear: tcgen05 precisions need dim %% 128 == 0");
  FBN_REQUIRE(aligned16(v) && aligned16(w), FBN_ERR_ALIGN, "bilinear: unaligned pointer");
  return FBN_OK;
}

extern "C" int fbn_bilinear_fwd(const float* v, const float* w, int type, int64_t batch, int fields, int dim, float* p, void* scratch,
                                size_t scratch_bytes, int precision, fbn_stream_t stream) {
  return fbn_bilinear_fwd_ld(v, w, type, batch, fields, dim, p, (int64_t)(fields * (fields - 1) / 2) * dim, scratch, scratch_bytes, precision,
                             stream);
}

extern "C" int fbn_bilinear_fwd_ld(const float* v, const float* w, int type, int64_t batch, int fields, int dim, float* p, int64_t ldp,
                                   void* scratch, size_t scratch_bytes, int precision, fbn_stream_t stream) {
  int rc = bil_check(v, w, type, batch, fields, dim, precision);
  if (rc) return rc;
  FBN_REQUIRE(p && scratch && scratch_bytes >= fbn_bilinear_scratch_bytes(batch, fields, dim, type), FBN_ERR_ARG,
              "fbn_bilinear_fwd: scratch too small");
  FBN_REQUIRE(ldp >= (int64_t)(fields * (fields - 1) / 2) * dim && ldp % 4 == 0 && aligned16(p), FBN_ERR_SHAPE, "fbn_bilinear_fwd: bad ldp");
  cudaStream_t st = (cudaStream_t)stream;
  BilScratch s = bil_carve(scratch, scratch_bytes, batch, fields, dim, type);
  rc = bil_transform(v, w, type, batch, fields, dim, s.T, precision, s, st);
  if (rc) return rc;
  const long long total = (long long)batch * (fields * (fields - 1) / 2) * (dim / 4);
  bil_pairs_fwd_kernel<<<ew_blocks(total), 256, 0, st>>>(v, (long long)fields * dim, s.T, type, batch, fields, dim, p, ldp);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_bilinear_bwd(const float* v, const float* w, const float* dp, int type, int64_t batch, int fields, int dim, float* dv,
                                float* dw, void* scratch, size_t scratch_bytes, int precision, fbn_stream_t stream) {
  return fbn_bilinear_bwd_ld(v, w, dp, (int64_t)(fields * (fields - 1) / 2) * dim, nullptr, 0, type, batch, fields, dim, dv, dw, scratch,
                             scratch_bytes, precision, stream);
}

extern "C" int fbn_bilinear_bwd_ld(const float* v, const float* w, const float* dp, int64_t lddp, const float* dv_init, int64_t ld_init,
                                   int type, int64_t batch, int fields, int dim, float* dv, float* dw, void* scratch, size_t scratch_bytes,
                                   int precision, fbn_stream_t stream) {
  int rc = bil_check(v, w, type, batch, fields, dim, precision);
  if (rc) return rc;
  FBN_REQUIRE(lddp % 4 == 0 && ld_init % 4 == 0 && aligned16(dp) && aligned16(dv_init), FBN_ERR_ALIGN, "fbn_bilinear_bwd: bad strides");
  FBN_REQUIRE(dp && dv && dw && scratch && scratch_bytes >= fbn_bilinear_scratch_bytes(batch, fields, dim, type), FBN_ERR_ARG,
              "fbn_bilinear_bwd: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int F = fields, Dm = dim;
  const long long B = batch;
  int nT, nW;
  bil_sizes(type, F, nT, nW);
  BilScratch s = bil_carve(scratch, scratch_bytes, batch, fields, dim, type);
  rc = bil_transform(v, w, type, batch, fields, dim, s.T, precision, s, st);     // recompute T (not saved by forward)
  if (rc) return rc;
  bil_pairs_bwd_kernel<<<ew_blocks(B * F * (Dm / 4)), 256, 0, st>>>(v, (long long)F * Dm, s.T, dp, lddp, dv_init, ld_init, type, B, F, Dm, dv,
                                                                     s.dT);
  FBN_CHECK_LAUNCH();
  // dv[src] += dT W^T ; dW = sum v_src^T dT
  GemmArgs d;
  d.N = Dm; d.K = Dm; d.ldb = Dm; d.b_t = 1; d.accumulate = 1;
  GemmArgs g;   // weight gradient, split-K over the contraction
  g.a_t = 1; g.b_t = 0; g.M = Dm; g.N = Dm; g.ldc = Dm;
  if (type == FBN_BILINEAR_ALL) {
    d.A = s.dT; d.lda = Dm; d.B = w; d.C = dv; d.ldc = Dm; d.M = B * F;
    rc = gemm(d, precision, s.tc, s.tc_bytes, st);
    if (rc) return rc;
    g.A = v; g.lda = Dm; g.B = s.dT; g.ldb = Dm; g.K = B * F;
    g.splits = (int)std::max<long long>(1, std::min<long long>(32, cdiv(g.K, 16)));
    g.C = s.partial; g.strideSplit = (long long)Dm * Dm;
    rc = gemm(g, precision, s.tc, s.tc_bytes, st);
    if (rc) return rc;
    return reduce_splits(s.partial, g.splits, Dm, Dm, (long long)Dm * Dm, ~0ull, dw, st);
  }
  const int P = F * (F - 1) / 2;
  g.K = B; g.lda = (long long)F * Dm; g.ldb = (long long)nT * Dm;
  g.splits = (int)std::max<long long>(1, std::min<long long>(32, cdiv(B, 16)));
  g.strideSplit = (long long)Dm * Dm;
  d.M = B; d.lda = (long long)nT * Dm; d.ldc = (long long)F * Dm;
  for (int t = 0; t < nT; ++t) {
    int src = t;
    if (type == FBN_BILINEAR_INTERACTION) {
      int i = 0, rem = t;
      while (rem >= F - 1 - i) { rem -= F - 1 - i; ++i; }
      src = i;
    }
    (void)P;
    d.A = s.dT + t * Dm; d.B = w + (long long)t * Dm * Dm; d.C = dv + src * Dm;
    rc = gemm(d, precision, s.tc, s.tc_bytes, st);
    if (rc) return rc;
    g.A = v + src * Dm; g.B = s.dT + t * Dm; g.C = s.partial;
    rc = gemm(g, precision, s.tc, s.tc_bytes, st);
    if (rc) return rc;
    rc = reduce_splits(s.partial, g.splits, Dm, Dm, (long long)Dm * Dm, ~0ull, dw + (long long)t * Dm * Dm, st);
    if (rc) return rc;
  }
  return FBN_OK;
}

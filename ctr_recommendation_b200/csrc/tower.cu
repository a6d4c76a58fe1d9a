// DNN tower glue around the GEMMs: BatchNorm1d statistics / apply / backward, ReLU, dropout,
// the 256->1 head with sigmoid, and the bilinear pair Hadamard products.
// Reference: src/model_fibinet.py:60-89 (pairs), :125-136,:197-199 (tower) and their autograd.
#include "common.cuh"
#include "tower.h"

namespace fbn {

// ------------------------------------------------------------------------------------------
// deterministic column reductions over the batch: partial[chunk][q][N]
// ------------------------------------------------------------------------------------------
enum { OP_SUM = 0, OP_SQDEV = 1, OP_BNBWD = 2, OP_HEADBWD = 3, OP_PROD2 = 4, OP_SHIFT2 = 5 };

struct ColArgs {
  const float* X; const float* Y; const float* Z;  // (B,N) operands (meaning depends on OP)
  const float* v0; const float* v1; const float* v2;  // per-column vectors
  const float* r0;                                  // per-row vector
  float scale;
  long long B; int N; long long rows_per_chunk;
  float* partial;
};

template <int OP>
__device__ __forceinline__ void col_terms(const ColArgs& a, long long r, int c, float4& q0, float4& q1, float4& q2) {
  const long long o = r * a.N + c;
  if (OP == OP_SUM) {
    q0 += ld4s(a.X + o);
  } else if (OP == OP_SQDEV) {
    const float4 d = ld4s(a.X + o) - ld4(a.v0 + c);
    q0 += d * d;
  } else if (OP == OP_SHIFT2) {   // sums of (x - pivot) and (x - pivot)^2, pivot = row 0 of the same column
    const float4 d = ld4s(a.X + o) - ld4(a.v0 + c);
    q0 += d;
    q1 += d * d;
  } else if (OP == OP_PROD2) {
    const float4 x = ld4s(a.X + o);
    q0 += x * ld4s(a.Y + o);
    q1 += x;
  } else {
    // X = dA (OP_BNBWD) ; Y = A (post relu/dropout) ; Z = H (pre-BN) ; v0 = mean ; v1 = rstd
    // OP_HEADBWD: dA = r0[r] * v2[c]  (dlogit x w3), q2 = dlogit * A
    const float4 A = ld4s(a.Y + o);
    float4 dA;
    if (OP == OP_HEADBWD) {
      const float dl = __ldg(a.r0 + r);
      dA = ld4(a.v2 + c) * dl;
      q2 += A * dl;
    } else {
      dA = ld4s(a.X + o);
    }
    float4 dY;
    dY.x = A.x > 0.f ? dA.x * a.scale : 0.f; dY.y = A.y > 0.f ? dA.y * a.scale : 0.f;
    dY.z = A.z > 0.f ? dA.z * a.scale : 0.f; dY.w = A.w > 0.f ? dA.w * a.scale : 0.f;
    const float4 xh = (ld4s(a.Z + o) - ld4(a.v0 + c)) * ld4(a.v1 + c);
    q0 += dY;
    q1 += dY * xh;
  }
}

template <int OP, int NQ>
__global__ void __launch_bounds__(256) colreduce_kernel(ColArgs a) {
  __shared__ float4 red[8][NQ][32];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + cg * 4;
  const long long r0 = (long long)blockIdx.y * a.rows_per_chunk, r1 = min(a.B, r0 + a.rows_per_chunk);
  float4 q0 = f4(0.f), q1 = f4(0.f), q2 = f4(0.f);
  if (c < a.N)
    for (long long r = r0 + rl; r < r1; r += 8) col_terms<OP>(a, r, c, q0, q1, q2);
  red[rl][0][cg] = q0;
  if (NQ > 1) red[rl][1 % NQ][cg] = q1;
  if (NQ > 2) red[rl][2 % NQ][cg] = q2;
  __syncthreads();
  if (rl < NQ && c < a.N) {
    float4 t = red[0][rl][cg];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][rl][cg];
    st4(a.partial + ((long long)blockIdx.y * NQ + rl) * a.N + c, t);
  }
}

// CTAs per SM of the column reductions (fbn_set_option("col_chunk_mult", m); partial buffers hold up to m = 16)
static int g_col_mult = 2;   // measured at B = 65536: 2 -> 4.135 ms / step, 4 -> 4.159, 8 -> 4.181
void set_col_chunk_mult(int m) { g_col_mult = std::max(1, std::min(m, 16)); }

int col_chunks(long long B, int N) {
  long long want = std::max<long long>(1, ((long long)g_col_mult * num_sms()) / std::max(1, N / 128));
  long long chunks = std::min<long long>(want, (B + 31) / 32);
  return (int)std::max<long long>(1, chunks);
}

template <int OP, int NQ>
static int launch_colreduce(ColArgs a, int chunks, cudaStream_t st) {
  a.rows_per_chunk = (a.B + chunks - 1) / chunks;
  dim3 grid((a.N + 127) / 128, chunks);
  colreduce_kernel<OP, NQ><<<grid, 256, 0, st>>>(a);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// out_q[c] = scale * sum_chunks partial[chunk][q][c].  One warp per output: lanes stride over the chunks and a
// fixed xor-tree combines them (double accumulation, run-to-run deterministic), so the latency is
// ceil(chunks/32) dependent loads instead of `chunks`.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) colfinal_kernel(const float* __restrict__ partial, int chunks, int N, int nq, float scale,
                                                       float* o0, float* o1, float* o2) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= N * nq) return;
  const int q = i / N, c = i % N;
  double t = 0.0;
  for (int k = lane; k < chunks; k += 32) t += (double)partial[((long long)k * nq + q) * N + c];
  t = warp_sum_d(t);
  float* o = q == 0 ? o0 : (q == 1 ? o1 : o2);
  if (lane == 0 && o) o[c] = (float)(t * scale);
}

static int launch_colfinal(const float* partial, int chunks, int N, int nq, float scale, float* o0, float* o1, float* o2,
                           cudaStream_t st) {
  colfinal_kernel<<<(N * nq + 7) / 8, 256, 0, st>>>(partial, chunks, N, nq, scale, o0, o1, o2);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

int colsum(const float* X, long long B, int N, float* partial, float* out, cudaStream_t st) {
  ColArgs a{}; a.X = X; a.B = B; a.N = N; a.partial = partial;
  const int ch = col_chunks(B, N);
  int rc = launch_colreduce<OP_SUM, 1>(a, ch, st);
  if (rc) return rc;
  return launch_colfinal(partial, ch, N, 1, 1.0f, out, nullptr, nullptr, st);
}

int colprod2(const float* X, const float* Y, long long B, int N, float* partial, float* out_xy, float* out_x, cudaStream_t st) {
  ColArgs a{}; a.X = X; a.Y = Y; a.B = B; a.N = N; a.partial = partial;
  const int ch = col_chunks(B, N);
  int rc = launch_colreduce<OP_PROD2, 2>(a, ch, st);
  if (rc) return rc;
  return launch_colfinal(partial, ch, N, 2, 1.0f, out_xy, out_x, nullptr, st);
}

// var finalize: rstd + running statistics update (momentum 0.1, unbiased variance), nn.BatchNorm1d
__global__ void __launch_bounds__(256) bn_var_final_kernel(const float* __restrict__ partial, int chunks, int N, long long B,
                                                           const float* mean, float* rstd, float* run_mean, float* run_var) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= N) return;
  double t = 0.0;
  for (int k = lane; k < chunks; k += 32) t += (double)partial[(long long)k * N + c];
  t = warp_sum_d(t);
  if (lane != 0) return;
  const float var = (float)(t / (double)B);
  rstd[c] = 1.0f / sqrtf(var + 1e-5f);
  if (run_mean) {
    const float unb = B > 1 ? (float)(t / (double)(B - 1)) : var;
    run_mean[c] = (1.0f - 0.1f) * run_mean[c] + 0.1f * mean[c];
    run_var[c] = (1.0f - 0.1f) * run_var[c] + 0.1f * unb;
  }
}

// single-pass batch statistics: with the pivot p_c = H[0][c] (a sample of the column, so |mean - p| ~ sigma and the
// subtraction below loses about one bit), mean = p + S1/B and var = S2/B - (S1/B)^2, finalised in fp64
__global__ void __launch_bounds__(256) bn_stats_final_kernel(const float* __restrict__ partial, int chunks, int N, long long B,
                                                             const float* __restrict__ pivot, float* mean, float* rstd,
                                                             float* run_mean, float* run_var) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= N) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = lane; k < chunks; k += 32) {
    s1 += (double)partial[((long long)k * 2 + 0) * N + c];
    s2 += (double)partial[((long long)k * 2 + 1) * N + c];
  }
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  if (lane != 0) return;
  const double m1 = s1 / (double)B;
  const double varb = fmax(s2 / (double)B - m1 * m1, 0.0);
  const float mu = (float)((double)pivot[c] + m1);
  mean[c] = mu;
  rstd[c] = 1.0f / sqrtf((float)varb + 1e-5f);
  if (run_mean) {
    const float unb = B > 1 ? (float)(varb * (double)B / (double)(B - 1)) : (float)varb;
    run_mean[c] = (1.0f - 0.1f) * run_mean[c] + 0.1f * mu;
    run_var[c] = (1.0f - 0.1f) * run_var[c] + 0.1f * unb;
  }
}

int bn_train_stats(const float* H, long long B, int N, float* partial, float* mean, float* rstd, float* run_mean, float* run_var,
                   cudaStream_t st) {
  ColArgs a{}; a.X = H; a.v0 = H; a.B = B; a.N = N; a.partial = partial;
  const int ch = col_chunks(B, N);
  int rc = launch_colreduce<OP_SHIFT2, 2>(a, ch, st);
  if (rc) return rc;
  bn_stats_final_kernel<<<(N + 7) / 8, 256, 0, st>>>(partial, ch, N, B, H, mean, rstd, run_mean, run_var);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

__global__ void bn_eval_stats_kernel(const float* run_mean, const float* run_var, int N, float* mean, float* rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  mean[c] = run_mean[c];
  rstd[c] = 1.0f / sqrtf(run_var[c] + 1e-5f);
}

int bn_eval_stats(const float* run_mean, const float* run_var, int N, float* mean, float* rstd, cudaStream_t st) {
  bn_eval_stats_kernel<<<(N + 127) / 128, 128, 0, st>>>(run_mean, run_var, N, mean, rstd);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// ------------------------------------------------------------------------------------------
// A = dropout(relu(bn(H)))  (elementwise, float4)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 bn_relu_drop4(float4 h, float4 mean, float4 rstd, float4 g, float4 b, const DropArgs& d,
                                                long long elem, const uint8_t* mask) {
  float4 y = (h - mean) * rstd * g + b;
  y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
  if (d.p > 0.f) {
    float k0, k1, k2, k3;
    if (mask) {
      const uchar4 m = *reinterpret_cast<const uchar4*>(mask + elem);
      k0 = m.x ? 1.f : 0.f; k1 = m.y ? 1.f : 0.f; k2 = m.z ? 1.f : 0.f; k3 = m.w ? 1.f : 0.f;
    } else {
      const uint64_t off = d.offset + (d.step_dev ? ((uint64_t)(uint32_t)__ldg(d.step_dev) << 36) : 0ull);
      const uint4 r = philox4x32(d.seed, off + (uint64_t)(elem >> 2), d.stream);
      k0 = u01(r.x) >= d.p ? 1.f : 0.f; k1 = u01(r.y) >= d.p ? 1.f : 0.f;
      k2 = u01(r.z) >= d.p ? 1.f : 0.f; k3 = u01(r.w) >= d.p ? 1.f : 0.f;
    }
    const float inv = 1.0f / (1.0f - d.p);
    // torch: out = input * mask / (1-p)  (mask * (1/(1-p)) then multiply)
    y.x = y.x * (k0 * inv); y.y = y.y * (k1 * inv); y.z = y.z * (k2 * inv); y.w = y.w * (k3 * inv);
  }
  return y;
}

__global__ void bn_act_kernel(const float* __restrict__ H, const float* __restrict__ mean, const float* __restrict__ rstd,
                              const float* __restrict__ g, const float* __restrict__ b, long long B, int N, DropArgs d,
                              float* __restrict__ A, PackDst pk) {
  const long long total4 = B * N / 4;
  float amax = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const int c = (int)(e % N);
    const float4 y = bn_relu_drop4(ld4s(H + e), ld4(mean + c), ld4(rstd + c), ld4(g + c), ld4(b + c), d, e, d.mask);
    st4(A + e, y);
    if (pk.mode) store_packed4(pk.base, pk.lo_off, pk.mode, e, y);     // pitch == N
    amax = fmaxf(amax, amax4(y));
  }
  if (pk.tail) f16x3_publish_amax(amax, pk.tail);
}

int bn_act(const float* H, const float* mean, const float* rstd, const float* g, const float* b, long long B, int N,
           const DropArgs& d, float* A, PackDst pk, cudaStream_t st) {
  const long long total4 = B * N / 4;
  int blocks = (int)std::min<long long>((total4 + 255) / 256, 8LL * num_sms());
  if (pk.tail) blocks = std::min(blocks, F16_AMAX_BLOCKS);
  bn_act_kernel<<<std::max(blocks, 1), 256, 0, st>>>(H, mean, rstd, g, b, B, N, d, A, pk);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// head: A2 = dropout(relu(bn(H2))); logit = A2 . w3 + b3; prob = sigmoid(logit)   (ref :131-136,:199)
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ H, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ g,
                                                       const float* __restrict__ b, const float* __restrict__ w3,
                                                       const float* __restrict__ b3, long long B, DropArgs d,
                                                       float* __restrict__ A, float* __restrict__ logit, float* __restrict__ prob) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int c0 = 4 * lane, c1 = 128 + 4 * lane;
  const float4 m0 = ld4(mean + c0), m1 = ld4(mean + c1), r0 = ld4(rstd + c0), r1 = ld4(rstd + c1);
  const float4 g0 = ld4(g + c0), g1 = ld4(g + c1), be0 = ld4(b + c0), be1 = ld4(b + c1);
  const float4 w0 = ld4(w3 + c0), w1 = ld4(w3 + c1);
  const float bias = __ldg(b3);
  for (long long r = warp; r < B; r += nwarps) {
    const long long e0 = r * H2 + c0, e1 = r * H2 + c1;
    const float4 a0 = bn_relu_drop4(ld4s(H + e0), m0, r0, g0, be0, d, e0, d.mask);
    const float4 a1 = bn_relu_drop4(ld4s(H + e1), m1, r1, g1, be1, d, e1, d.mask);
    st4(A + e0, a0);
    st4(A + e1, a1);
    const float z = warp_sum(hsum4(a0 * w0) + hsum4(a1 * w1)) + bias;
    if (lane == 0) {
      logit[r] = z;
      prob[r] = sigmoidf_(z);
    }
  }
}

int head_fwd(const float* H, const float* mean, const float* rstd, const float* g, const float* b, const float* w3, const float* b3,
             long long B, const DropArgs& d, float* A, float* logit, float* prob, cudaStream_t st) {
  int blocks = (int)std::min<long long>((B + 7) / 8, 8LL * num_sms());
  head_fwd_kernel<<<std::max(blocks, 1), 256, 0, st>>>(H, mean, rstd, g, b, w3, b3, B, d, A, logit, prob);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// dlogit = dprob * p * (1-p)  (sigmoid backward) + per-block double partial sums for db3 (contiguous chunk per block, fixed order)
__global__ void __launch_bounds__(256) head_dlogit_kernel(const float* __restrict__ dprob, const float* __restrict__ prob, long long B,
                                                          long long per, float* __restrict__ dlogit, double* __restrict__ partial) {
  __shared__ double s[256];
  const long long b0 = (long long)blockIdx.x * per, b1 = min(B, b0 + per);
  double t = 0.0;
  for (long long i = b0 + threadIdx.x; i < b1; i += 256) {
    const float p = prob[i];
    const float d = dprob[i] * ((1.0f - p) * p);
    dlogit[i] = d;
    t += (double)d;
  }
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

__global__ void sum1_kernel(const double* __restrict__ x, int n, float* out) {  // single block over the per-block partials, fixed order
  __shared__ double s[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += x[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)s[0];
}

// Layer-2 backward statistics straight from dlogit (dA2 = dlogit x w3 is never materialised):
//   dgamma2 = sum dY*xhat, dbeta2 = sum dY, dw3 = sum dlogit*A2, db3 = sum dlogit
int head_bwd_stats(const float* dprob, const float* prob, const float* A2, const float* Hd2, const float* mean, const float* rstd,
                   const float* w3, long long B, float scale, float* partial, float* dlogit, float* dgamma, float* dbeta, float* dw3,
                   float* db3, cudaStream_t st) {
  const int blocks = (int)std::max<long long>(1, std::min<long long>((B + 1023) / 1024, 2LL * num_sms()));
  double* dpart = reinterpret_cast<double*>(partial);        // <= 296 doubles; consumed before the column reduction reuses `partial`
  head_dlogit_kernel<<<blocks, 256, 0, st>>>(dprob, prob, B, (B + blocks - 1) / blocks, dlogit, dpart);
  FBN_CHECK_LAUNCH();
  sum1_kernel<<<1, 256, 0, st>>>(dpart, blocks, db3);
  FBN_CHECK_LAUNCH();
  ColArgs a{}; a.Y = A2; a.Z = Hd2; a.v0 = mean; a.v1 = rstd; a.v2 = w3; a.r0 = dlogit; a.scale = scale; a.B = B; a.N = H2;
  a.partial = partial;
  const int ch = col_chunks(B, H2);
  int rc = launch_colreduce<OP_HEADBWD, 3>(a, ch, st);
  if (rc) return rc;
  return launch_colfinal(partial, ch, H2, 3, 1.0f, dbeta, dgamma, dw3, st);
}

int bn_bwd_stats(const float* dA, const float* A, const float* Hd, const float* mean, const float* rstd, long long B, int N,
                 float scale, float* partial, float* dgamma, float* dbeta, cudaStream_t st) {
  ColArgs a{}; a.X = dA; a.Y = A; a.Z = Hd; a.v0 = mean; a.v1 = rstd; a.scale = scale; a.B = B; a.N = N; a.partial = partial;
  const int ch = col_chunks(B, N);
  int rc = launch_colreduce<OP_BNBWD, 2>(a, ch, st);
  if (rc) return rc;
  return launch_colfinal(partial, ch, N, 2, 1.0f, dbeta, dgamma, nullptr, st);
}

// dH = gamma*rstd * (dY - dbeta/B - xhat*dgamma/B)  (train)   |   dH = dY*gamma*rstd  (eval)
// dY is recomputed from dA (or dlogit x w3) and the relu/dropout mask implied by A > 0.
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dA, const float* __restrict__ dlogit, const float* __restrict__ w3,
                                    const float* __restrict__ A, const float* __restrict__ Hd, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ g, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, long long B, int N, float scale, int train,
                                    float* __restrict__ dH, PackDst pk) {
  const long long total4 = B * N / 4;
  const float invB = 1.0f / (float)B;
  float amax = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const int c = (int)(e % N);
    const long long r = e / N;
    const float4 a = ld4s(A + e);
    const float4 da = dlogit ? ld4(w3 + c) * __ldg(dlogit + r) : ld4s(dA + e);
    float4 dY;
    dY.x = a.x > 0.f ? da.x * scale : 0.f; dY.y = a.y > 0.f ? da.y * scale : 0.f;
    dY.z = a.z > 0.f ? da.z * scale : 0.f; dY.w = a.w > 0.f ? da.w * scale : 0.f;
    const float4 rs = ld4(rstd + c);
    const float4 gr = ld4(g + c) * rs;
    float4 out;
    if (train) {
      const float4 xh = (ld4s(Hd + e) - ld4(mean + c)) * rs;
      out = gr * (dY - ld4(dbeta + c) * invB - xh * (ld4(dgamma + c) * invB));
    } else {
      out = gr * dY;
    }
    st4(dH + e, out);
    if (pk.mode) store_packed4(pk.base, pk.lo_off, pk.mode, e, out);
    amax = fmaxf(amax, amax4(out));
  }
  if (pk.tail) f16x3_publish_amax(amax, pk.tail);
}

int bn_bwd_apply(const float* dA, const float* dlogit, const float* w3, const float* A, const float* Hd, const float* mean,
                 const float* rstd, const float* g, const float* dgamma, const float* dbeta, long long B, int N, float scale,
                 int train, float* dH, PackDst pk, cudaStream_t st) {
  const long long total4 = B * N / 4;
  int blocks = (int)std::min<long long>((total4 + 255) / 256, 8LL * num_sms());
  if (pk.tail) blocks = std::min(blocks, F16_AMAX_BLOCKS);
  bn_bwd_apply_kernel<<<std::max(blocks, 1), 256, 0, st>>>(dA, dlogit, w3, A, Hd, mean, rstd, g, dgamma, dbeta, B, N, scale, train,
                                                           dH, pk);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// ------------------------------------------------------------------------------------------
// bilinear pair products.  V lives in C (blocks 1..5), T = transformed fields (B,nT,128).
//   ALL        : T_t = V_{t+2} W          p_ij = V_i  * T_{j-2}      (ref :72,79)
//   EACH       : T_t = V_{t+1} W_{t+1}    p_ij = T_{i-1} * V_j       (ref :85-86)
//   INTERACTION: T_q = V_i W_(i,j)        p_ij = T_q * V_j           (extension)
// Pairs with i = 0 are identically zero (V_0 = 0) and their C blocks are never written.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pair_block(int i, int j) { return NF + i * (2 * NF - i - 1) / 2 + (j - i - 1); }
__host__ __device__ constexpr int active_q(int i, int j) {  // index among pairs with i >= 1
  return (i - 1) * (2 * NA - (i - 1) - 1) / 2 + (j - i - 1);
}

template <int TYPE>
__global__ void __launch_bounds__(256) bilinear_pairs_fwd_kernel(float* __restrict__ C, const float* __restrict__ T, long long B, PackDst pk) {
  constexpr int nT = TYPE == FBN_BILINEAR_INTERACTION ? 10 : 4;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp; b < B; b += nwarps) {
    float* crow = C + b * K1 + 4 * lane;
    float4 v[NF], t[nT];
#pragma unroll
    for (int f = 1; f < NF; ++f) v[f] = *reinterpret_cast<const float4*>(crow + f * D);
#pragma unroll
    for (int k = 0; k < nT; ++k) t[k] = ld4s(T + b * (nT * D) + k * D + 4 * lane);
#pragma unroll
    for (int i = 1; i < NF; ++i)
#pragma unroll
      for (int j = i + 1; j < NF; ++j) {
        float4 p;
        if (TYPE == FBN_BILINEAR_ALL) p = v[i] * t[j - 2];
        else if (TYPE == FBN_BILINEAR_EACH) p = t[i - 1] * v[j];
        else p = t[active_q(i, j)] * v[j];
        // with a tcgen05 precision only the GEMMs read the pair blocks, and they read the packed copy
        if (pk.mode) store_packed4(pk.base, pk.lo_off, pk.mode, b * K1 + pair_block(i, j) * D + 4 * lane, p);
        else st4(crow + pair_block(i, j) * D, p);
      }
  }
}

int bilinear_pairs_fwd(int type, float* C, const float* T, long long B, PackDst pk, cudaStream_t st) {
  int blocks = (int)std::min<long long>((B + 7) / 8, 8LL * num_sms());
  blocks = std::max(blocks, 1);
  if (type == FBN_BILINEAR_ALL) bilinear_pairs_fwd_kernel<FBN_BILINEAR_ALL><<<blocks, 256, 0, st>>>(C, T, B, pk);
  else if (type == FBN_BILINEAR_EACH) bilinear_pairs_fwd_kernel<FBN_BILINEAR_EACH><<<blocks, 256, 0, st>>>(C, T, B, pk);
  else bilinear_pairs_fwd_kernel<FBN_BILINEAR_INTERACTION><<<blocks, 256, 0, st>>>(C, T, B, pk);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// FBN_PREC_F16X3: the whole live MLP-input row (fields 1..5 and the 10 live pairs) as fp16 hi|lo under ONE scale.  The scale needs
// the maximum over values this very stage computes, so it runs twice: PACK = false only reduces max |v|, max |pair| per block
// (reads 4.5 KB / sample, writes nothing), PACK = true recomputes the products and writes the 15 split blocks (7.7 KB / sample).
// The fp32 pair blocks are never materialised (as in the other tcgen05 modes), and there is no separate pass over C.
template <int TYPE, bool PACK>
__global__ void __launch_bounds__(256) bilinear_pairs_mlp16_kernel(const float* __restrict__ C, const float* __restrict__ T, long long B,
                                                                   __half* __restrict__ dst, long long lo_off, float* __restrict__ tail) {
  constexpr int nT = TYPE == FBN_BILINEAR_INTERACTION ? 10 : 4;
  float sc = 1.f;
  if (PACK) sc = f16x3_block_scale(tail, -1);
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float amax = 0.f;
  for (long long b = warp; b < B; b += nwarps) {
    const float* crow = C + b * K1 + 4 * lane;
    float4 v[NF], t[nT];
#pragma unroll
    for (int f = 1; f < NF; ++f) v[f] = ld4s(crow + f * D);
#pragma unroll
    for (int k = 0; k < nT; ++k) t[k] = ld4s(T + b * (nT * D) + k * D + 4 * lane);
    const long long e0 = b * K1 + 4 * lane;
#pragma unroll
    for (int f = 1; f < NF; ++f) {
      if (PACK) store_f16x3_4(dst, lo_off, e0 + f * D, v[f], sc);
      else amax = fmaxf(amax, amax4(v[f]));
    }
#pragma unroll
    for (int i = 1; i < NF; ++i)
#pragma unroll
      for (int j = i + 1; j < NF; ++j) {
        float4 p;
        if (TYPE == FBN_BILINEAR_ALL) p = v[i] * t[j - 2];
        else if (TYPE == FBN_BILINEAR_EACH) p = t[i - 1] * v[j];
        else p = t[active_q(i, j)] * v[j];
        if (PACK) store_f16x3_4(dst, lo_off, e0 + pair_block(i, j) * D, p, sc);
        else amax = fmaxf(amax, amax4(p));
      }
  }
  if (!PACK) f16x3_publish_amax(amax, tail);
}

template <int TYPE>
static int launch_pairs_mlp16(const float* C, const float* T, long long B, void* dst16, long long lo_off, float* tail, cudaStream_t st) {
  int blocks = (int)std::min<long long>((B + 7) / 8, 8LL * num_sms());
  blocks = std::max(blocks, 1);
  bilinear_pairs_mlp16_kernel<TYPE, false><<<std::min(blocks, F16_AMAX_BLOCKS), 256, 0, st>>>(C, T, B, nullptr, 0, tail);
  FBN_CHECK_LAUNCH();
  bilinear_pairs_mlp16_kernel<TYPE, true><<<blocks, 256, 0, st>>>(C, T, B, static_cast<__half*>(dst16), lo_off, tail);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

int bilinear_pairs_mlp16(int type, const float* C, const float* T, long long B, void* dst16, long long lo_off, float* tail, cudaStream_t st) {
  if (type == FBN_BILINEAR_ALL) return launch_pairs_mlp16<FBN_BILINEAR_ALL>(C, T, B, dst16, lo_off, tail, st);
  if (type == FBN_BILINEAR_EACH) return launch_pairs_mlp16<FBN_BILINEAR_EACH>(C, T, B, dst16, lo_off, tail, st);
  return launch_pairs_mlp16<FBN_BILINEAR_INTERACTION>(C, T, B, dst16, lo_off, tail, st);
}

// dT_t = sum_q dP_q * (other operand) ; dV_f = dC_f + sum_q dP_q * T  (the W^T term is added by a GEMM)
template <int TYPE>
__global__ void __launch_bounds__(256) bilinear_pairs_bwd_kernel(const float* __restrict__ C, const float* __restrict__ T,
                                                                 const float* __restrict__ dC, long long B, float* __restrict__ dT,
                                                                 float* __restrict__ dV, PackDst pk) {
  constexpr int nT = TYPE == FBN_BILINEAR_INTERACTION ? 10 : 4;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp; b < B; b += nwarps) {
    const float* crow = C + b * K1 + 4 * lane;
    const float* drow = dC + b * K1 + 4 * lane;
    float4 v[NF], t[nT], dv[NF], dt[nT];
#pragma unroll
    for (int f = 1; f < NF; ++f) {
      v[f] = ld4s(crow + f * D);
      dv[f] = ld4s(drow + f * D);
    }
#pragma unroll
    for (int k = 0; k < nT; ++k) {
      t[k] = ld4s(T + b * (nT * D) + k * D + 4 * lane);
      dt[k] = f4(0.f);
    }
#pragma unroll
    for (int i = 1; i < NF; ++i)
#pragma unroll
      for (int j = i + 1; j < NF; ++j) {
        const float4 dp = ld4s(drow + pair_block(i, j) * D);
        if (TYPE == FBN_BILINEAR_ALL) {
          dv[i] += dp * t[j - 2];
          dt[j - 2] += dp * v[i];
        } else if (TYPE == FBN_BILINEAR_EACH) {
          dv[j] += dp * t[i - 1];
          dt[i - 1] += dp * v[j];
        } else {
          dv[j] += dp * t[active_q(i, j)];
          dt[active_q(i, j)] += dp * v[j];
        }
      }
#pragma unroll
    for (int f = 1; f < NF; ++f) st4(dV + b * (NA * D) + (f - 1) * D + 4 * lane, dv[f]);
#pragma unroll
    for (int k = 0; k < nT; ++k) {
      if (pk.mode) store_packed4(pk.base, pk.lo_off, pk.mode, b * (nT * D) + k * D + 4 * lane, dt[k]);   // only GEMMs read dT
      else st4(dT + b * (nT * D) + k * D + 4 * lane, dt[k]);
    }
  }
}

int bilinear_pairs_bwd(int type, const float* C, const float* T, const float* dC, long long B, float* dT, float* dV, PackDst pk,
                       cudaStream_t st) {
  int blocks = (int)std::min<long long>((B + 7) / 8, 8LL * num_sms());
  blocks = std::max(blocks, 1);
  if (type == FBN_BILINEAR_ALL) bilinear_pairs_bwd_kernel<FBN_BILINEAR_ALL><<<blocks, 256, 0, st>>>(C, T, dC, B, dT, dV, pk);
  else if (type == FBN_BILINEAR_EACH) bilinear_pairs_bwd_kernel<FBN_BILINEAR_EACH><<<blocks, 256, 0, st>>>(C, T, dC, B, dT, dV, pk);
  else bilinear_pairs_bwd_kernel<FBN_BILINEAR_INTERACTION><<<blocks, 256, 0, st>>>(C, T, dC, B, dT, dV, pk);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// reduce split-K partials of a (M,N) weight gradient; columns in inactive 128-blocks are written as 0
__global__ void reduce_splits_kernel(const float* __restrict__ partial, int parts, long long M, long long N, long long part_stride,
                                     unsigned long long nmask, float* __restrict__ out) {
  const long long total4 = M * N / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const long long n = e % N;
    float4 t = f4(0.f);
    if (nmask == ~0ull || ((nmask >> (n / 128)) & 1ull))
      for (int p = 0; p < parts; ++p) t += ld4s(partial + (long long)p * part_stride + e);
    st4(out + e, t);
  }
}

int reduce_splits(const float* partial, int parts, long long M, long long N, long long part_stride, unsigned long long nmask, float* out,
                  cudaStream_t st) {
  const long long total4 = M * N / 4;
  int blocks = (int)std::min<long long>((total4 + 255) / 256, 8LL * num_sms());
  reduce_splits_kernel<<<std::max(blocks, 1), 256, 0, st>>>(partial, parts, M, N, part_stride, nmask, out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

}  // namespace fbn

// N1 / A1 / A2: global gradient norm, clip coefficient, dense-exact Adam over the embedding table and
// flat Adam over the dense parameters.
// Reference: torch.nn.utils.clip_grad_norm_(.., 10.0) + torch.optim.Adam(lr, weight_decay) as invoked at
// src/train_fibinet.py:78,119,121 (single-tensor Adam math of torch/optim/adam.py, L2 decay folded
// into the gradient, bias corrections from the scheduler-cycled beta1).
#include <stdlib.h>

#include "common.cuh"
#include "tower.h"

namespace fbn {

constexpr int SQ_THREADS = 256;

__global__ void __launch_bounds__(SQ_THREADS) sumsq_kernel(const float* __restrict__ x, long long n, long long per_block,
                                                           float* __restrict__ partial) {
  __shared__ float s[SQ_THREADS];
  const long long b0 = (long long)blockIdx.x * per_block, b1 = min(n, b0 + per_block);
  float t = 0.f;
  for (long long i = b0 + threadIdx.x * 4; i < b1; i += SQ_THREADS * 4) {
    if (i + 3 < b1) {
      const float4 v = ld4s(x + i);
      t += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    } else {
      for (long long j = i; j < b1; ++j) t += x[j] * x[j];
    }
  }
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = SQ_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

__global__ void sumsq_final_kernel(const float* __restrict__ partial, int n, float* out) {
  __shared__ double s[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += (double)partial[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)s[0];
}

int sumsq_partial_count(long long n) {
  // chunk boundaries are multiples of 4 floats so that the float4 loads stay aligned
  long long blocks = std::min<long long>(std::max<long long>(1, n / 4096), 1024);
  return (int)blocks;
}

int sumsq(const float* x, long long n, float* partial, float* out, cudaStream_t st) {
  const int blocks = sumsq_partial_count(n);
  long long per = cdiv(cdiv(n, blocks), 4) * 4;
  sumsq_kernel<<<blocks, SQ_THREADS, 0, st>>>(x, n, per, partial);
  FBN_CHECK_LAUNCH();
  sumsq_final_kernel<<<1, 256, 0, st>>>(partial, blocks, out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

__global__ void clip_coef_kernel(const float* __restrict__ sums, int n, float max_norm, float* out) {
  double t = 0.0;
  for (int i = 0; i < n; ++i) t += (double)sums[i];
  const float total = (float)sqrt(t);
  out[0] = total;
  out[1] = fminf(1.0f, max_norm / (total + 1e-6f));  // clip_grad.py: clamp(max_norm/(total+1e-6), max=1)
}

struct AdamHyper { float lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, omb1, omb2, decay_mul; };   // decay_mul != 0: AdamW

__device__ __forceinline__ AdamHyper resolve_hyper(const AdamHyper& h, const float* dev) {
  if (!dev) return h;
  AdamHyper r;
  r.lr = dev[0]; r.beta1 = dev[1]; r.beta2 = dev[2]; r.eps = dev[3]; r.wd = dev[4]; r.step_size = dev[5]; r.bc2_sqrt = dev[6];
  r.omb1 = dev[8]; r.omb2 = dev[9]; r.decay_mul = dev[10];
  return r;
}

__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, const AdamHyper& h) {
  if (h.decay_mul != 0.f) p *= h.decay_mul;          // AdamW: param.mul_(1 - lr * weight_decay)
  else g = fmaf(h.wd, p, g);                         // Adam: grad.add(param, alpha=weight_decay)
  m = m + h.omb1 * (g - m);                          // exp_avg.lerp_(grad, 1-beta1)      (1-beta evaluated in double by torch)
  v = v * h.beta2 + h.omb2 * (g * g);                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps; // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - h.step_size * (m / denom);                 // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// One float4 per thread, 32 threads per table row; untouched rows still get g = wd*p (dense-exact).
__global__ void __launch_bounds__(256) adam_table_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                         const float* __restrict__ grad, const int32_t* __restrict__ touched,
                                                         long long rows, const float* __restrict__ clip, AdamHyper hv,
                                                         const float* __restrict__ hdev) {
  const AdamHyper h = resolve_hyper(hv, hdev);
  const float coef = clip ? clip[1] : 1.0f;
  const long long total4 = rows * (D / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i >> 5;
    float4 g = f4(0.f);
    if (touched == nullptr || touched[r] > 0) g = ld4s(grad + i * 4) * coef;
    float4 pp = ld4s(p + i * 4), mm = ld4s(m + i * 4), vv = ld4s(v + i * 4);
    adam1(pp.x, mm.x, vv.x, g.x, h); adam1(pp.y, mm.y, vv.y, g.y, h);
    adam1(pp.z, mm.z, vv.z, g.z, h); adam1(pp.w, mm.w, vv.w, g.w, h);
    st4(p + i * 4, pp); st4(m + i * 4, mm); st4(v + i * 4, vv);
  }
}

__global__ void __launch_bounds__(256) adam_dense_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                         const float* __restrict__ grad, long long n, const float* __restrict__ clip,
                                                         AdamHyper hv, const float* __restrict__ hdev) {
  const AdamHyper h = resolve_hyper(hv, hdev);
  const float coef = clip ? clip[1] : 1.0f;
  const long long total4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = ld4s(grad + i * 4) * coef;
    float4 pp = ld4s(p + i * 4), mm = ld4s(m + i * 4), vv = ld4s(v + i * 4);
    adam1(pp.x, mm.x, vv.x, g.x, h); adam1(pp.y, mm.y, vv.y, g.y, h);
    adam1(pp.z, mm.z, vv.z, g.z, h); adam1(pp.w, mm.w, vv.w, g.w, h);
    st4(p + i * 4, pp); st4(m + i * 4, mm); st4(v + i * 4, vv);
  }
}

// ---- Adagrad (north_star (2): "scatter-add of sparse row gradients fused into the Adam/Adagrad row update") -----------------
// torch.optim.Adagrad, single-tensor path (torch/optim/adagrad.py): grad += wd * p ; clr = lr / (1 + (step - 1) * lr_decay) ;
// state_sum += grad * grad ; p -= clr * grad / (sqrt(state_sum) + eps).  The reference itself builds Adam (src/train_fibinet.py:78):
// extension, pinned to torch.optim.Adagrad on CPU (tests/test_oracle_optim.py).  hyper_dev = {clr, -, -, eps, wd, ...}.
struct AdagradHyper { float clr, eps, wd; };

__device__ __forceinline__ AdagradHyper resolve_adagrad(const AdagradHyper& h, const float* dev) {
  if (!dev) return h;
  AdagradHyper r;
  r.clr = dev[0]; r.eps = dev[3]; r.wd = dev[4];
  return r;
}

__device__ __forceinline__ void adagrad1(float& p, float& ssum, float g, const AdagradHyper& h) {
  g = fmaf(h.wd, p, g);                        // grad.add(param, alpha=weight_decay)  (wd == 0: g unchanged)
  ssum = fmaf(g, g, ssum);                     // state_sum.addcmul_(grad, grad, value=1)
  const float std_ = sqrtf(ssum) + h.eps;      // state_sum.sqrt().add_(eps)
  p = p - h.clr * (g / std_);                  // param.addcdiv_(grad, std, value=-clr)
}

// Row update over the table: a touched row takes its segment-summed gradient (x clip coefficient); an untouched row has
// g = wd * p -- with wd == 0 that is the identity (0 / (sqrt(sum) + eps) = 0), so the row is not even read: the update then
// moves 4 streams x 512 B per TOUCHED row only (the sparse row update), otherwise it is dense-exact like the Adam kernel.
__global__ void __launch_bounds__(256) adagrad_table_kernel(float* __restrict__ p, float* __restrict__ ssum, const float* __restrict__ grad,
                                                            const int32_t* __restrict__ touched, long long rows,
                                                            const float* __restrict__ clip, AdagradHyper hv, const float* __restrict__ hdev) {
  const AdagradHyper h = resolve_adagrad(hv, hdev);
  const float coef = clip ? clip[1] : 1.0f;
  const long long total4 = rows * (D / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i >> 5;
    const bool hit = touched == nullptr || touched[r] > 0;
    if (!hit && h.wd == 0.f) continue;
    const float4 g = hit ? ld4s(grad + i * 4) * coef : f4(0.f);
    float4 pp = ld4s(p + i * 4), ss = ld4s(ssum + i * 4);
    adagrad1(pp.x, ss.x, g.x, h); adagrad1(pp.y, ss.y, g.y, h); adagrad1(pp.z, ss.z, g.z, h); adagrad1(pp.w, ss.w, g.w, h);
    st4(p + i * 4, pp); st4(ssum + i * 4, ss);
  }
}

__global__ void __launch_bounds__(256) adagrad_dense_kernel(float* __restrict__ p, float* __restrict__ ssum, const float* __restrict__ grad,
                                                            long long n, const float* __restrict__ clip, AdagradHyper hv,
                                                            const float* __restrict__ hdev) {
  const AdagradHyper h = resolve_adagrad(hv, hdev);
  const float coef = clip ? clip[1] : 1.0f;
  const long long total4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = ld4s(grad + i * 4) * coef;
    float4 pp = ld4s(p + i * 4), ss = ld4s(ssum + i * 4);
    adagrad1(pp.x, ss.x, g.x, h); adagrad1(pp.y, ss.y, g.y, h); adagrad1(pp.z, ss.z, g.z, h); adagrad1(pp.w, ss.w, g.w, h);
    st4(p + i * 4, pp); st4(ssum + i * 4, ss);
  }
}

// 1 - beta as torch computes it: in double, from the decimal literal the user wrote.  An fp32 beta is the rounding of such a
// literal; the shortest decimal that rounds to it (<= 9 digits) recovers the double, e.g. 0.999f -> 0.999 -> 1e-3.
float one_minus(float beta) {
  char buf[32];
  for (int prec = 1; prec <= 9; ++prec) {
    snprintf(buf, sizeof(buf), "%.*g", prec, (double)beta);
    if ((float)atof(buf) == beta) return (float)(1.0 - atof(buf));
  }
  return 1.0f - beta;
}

static AdamHyper make_hyper(const fbn_adam_t& a) {
  AdamHyper h;
  h.lr = a.lr; h.beta1 = a.beta1; h.beta2 = a.beta2; h.eps = a.eps; h.wd = a.weight_decay;
  h.omb1 = a.one_minus_beta1 > 0.f ? a.one_minus_beta1 : one_minus(a.beta1);
  h.omb2 = a.one_minus_beta2 > 0.f ? a.one_minus_beta2 : one_minus(a.beta2);
  // bias corrections 1 - beta^t: the fp32 rounding of beta (0.999f = 0.99900001) would put a 1.3e-5 relative error into
  // 1 - beta2^t at t = 1 (cancellation), i.e. 6e-6 into every step; 1 - (1 - beta), with 1 - beta as torch evaluates it, does not
  const double b1 = 1.0 - (double)h.omb1, b2 = 1.0 - (double)h.omb2;
  const double bc1 = 1.0 - pow(b1, (double)a.step);
  const double bc2 = 1.0 - pow(b2, (double)a.step);
  h.step_size = (float)((double)a.lr / bc1);
  h.bc2_sqrt = (float)sqrt(bc2);
  h.decay_mul = a.decoupled ? (float)(1.0 - (double)a.lr * (double)a.weight_decay) : 0.f;
  return h;
}

// OneCycleLR (cos, two phases, cycle_momentum) evaluated on the device from a step counter, so that a
// whole training step can be replayed from a CUDA graph: hyper[0..6] as AdamHyper, counter += 1.
__global__ void onecycle_hyper_kernel(int* step_counter, int total_steps, float max_lr, float pct_start, float div_factor,
                                      float final_div, float base_m, float max_m, float beta2, float eps, float wd, float omb2,
                                      float* hyper) {
  const int s = *step_counter;  // 0-based scheduler step == number of optimizer steps already taken
  const double init = (double)max_lr / div_factor, minlr = init / final_div;
  const double e1 = (double)pct_start * total_steps - 1.0, e2 = total_steps - 1.0;
  double lr, b1;
  const double PI = 3.14159265358979323846;
  if ((double)s <= e1) {
    const double pct = (double)s / e1;
    const double c = cos(PI * pct) + 1.0;
    lr = max_lr + (init - max_lr) / 2.0 * c;
    b1 = base_m + (max_m - base_m) / 2.0 * c;
  } else {
    const double pct = ((double)s - e1) / (e2 - e1);
    const double c = cos(PI * pct) + 1.0;
    lr = minlr + (max_lr - minlr) / 2.0 * c;
    b1 = max_m + (base_m - max_m) / 2.0 * c;
  }
  const int t = s + 1;
  hyper[0] = (float)lr; hyper[1] = (float)b1; hyper[2] = beta2; hyper[3] = eps; hyper[4] = wd;
  hyper[5] = (float)(lr / (1.0 - pow(b1, (double)t)));
  hyper[6] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
  hyper[7] = (float)t;
  hyper[8] = (float)(1.0 - b1);
  hyper[9] = omb2;
  hyper[10] = 0.f;   // coupled L2 decay (the reference's torch.optim.Adam)
  *step_counter = t;
}

}  // namespace fbn

using namespace fbn;

extern "C" int fbn_clip_coef(const float* sumsq_in, int n, float max_norm, float* out, fbn_stream_t stream) {
  FBN_REQUIRE(sumsq_in && out && n > 0, FBN_ERR_ARG, "fbn_clip_coef: bad arguments");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq_in, n, max_norm, out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_adam_table(float* p, float* m, float* v, const float* grad, const int32_t* row_touched, int64_t rows,
                              const float* clip, const fbn_adam_t* h, const float* hyper_dev, fbn_stream_t stream) {
  FBN_REQUIRE(p && m && v && grad && (h || hyper_dev), FBN_ERR_ARG, "fbn_adam_table: null pointer");
  FBN_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v) && aligned16(grad), FBN_ERR_ALIGN, "fbn_adam_table: unaligned pointer");
  AdamHyper hv{};
  if (h) hv = make_hyper(*h);
  const long long total4 = rows * (D / 4);
  int blocks = (int)std::min<long long>(cdiv(total4, 256), 16LL * num_sms());
  adam_table_kernel<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(p, m, v, grad, row_touched, rows, clip, hv, hyper_dev);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_adam_dense(float* p, float* m, float* v, const float* grad, int64_t n, const float* clip, const fbn_adam_t* h,
                              const float* hyper_dev, fbn_stream_t stream) {
  FBN_REQUIRE(p && m && v && grad && (h || hyper_dev), FBN_ERR_ARG, "fbn_adam_dense: null pointer");
  FBN_REQUIRE(n % 4 == 0, FBN_ERR_SHAPE, "fbn_adam_dense: n must be a multiple of 4 (pad the flat buffer)");
  FBN_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v) && aligned16(grad), FBN_ERR_ALIGN, "fbn_adam_dense: unaligned pointer");
  AdamHyper hv{};
  if (h) hv = make_hyper(*h);
  int blocks = (int)std::min<long long>(cdiv(n / 4, 256), 16LL * num_sms());
  adam_dense_kernel<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(p, m, v, grad, n, clip, hv, hyper_dev);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_onecycle_hyper(int32_t* step_counter, int total_steps, float max_lr, float pct_start, float div_factor,
                                  float final_div_factor, float base_momentum, float max_momentum, float beta2, float eps,
                                  float weight_decay, float* hyper_dev, fbn_stream_t stream) {
  FBN_REQUIRE(step_counter && hyper_dev && total_steps > 1, FBN_ERR_ARG, "fbn_onecycle_hyper: bad arguments");
  onecycle_hyper_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter, total_steps, max_lr, pct_start, div_factor, final_div_factor,
                                                            base_momentum, max_momentum, beta2, eps, weight_decay,
                                                            one_minus(beta2), hyper_dev);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" size_t fbn_sumsq_partial_floats(int64_t n) { return (size_t)sumsq_partial_count(n); }

extern "C" int fbn_sumsq(const float* x, int64_t n, float* partial, float* out, fbn_stream_t stream) {
  FBN_REQUIRE(x && out && partial, FBN_ERR_ARG, "fbn_sumsq: null pointer");
  FBN_REQUIRE(aligned16(x), FBN_ERR_ALIGN, "fbn_sumsq: unaligned pointer");
  return sumsq(x, n, partial, out, (cudaStream_t)stream);
}

static AdagradHyper make_adagrad(const fbn_adagrad_t& a) {
  AdagradHyper h;
  h.clr = (float)((double)a.lr / (1.0 + (double)(a.step - 1) * (double)a.lr_decay));
  h.eps = a.eps; h.wd = a.weight_decay;
  return h;
}

extern "C" int fbn_adagrad_table(float* p, float* state_sum, const float* grad, const int32_t* row_touched, int64_t rows,
                                 const float* clip, const fbn_adagrad_t* h, const float* hyper_dev, fbn_stream_t stream) {
  FBN_REQUIRE(p && state_sum && grad && (h || hyper_dev), FBN_ERR_ARG, "fbn_adagrad_table: null pointer");
  FBN_REQUIRE(aligned16(p) && aligned16(state_sum) && aligned16(grad), FBN_ERR_ALIGN, "fbn_adagrad_table: unaligned pointer");
  AdagradHyper hv{};
  if (h) hv = make_adagrad(*h);
  const long long total4 = rows * (D / 4);
  int blocks = (int)std::min<long long>(cdiv(total4, 256), 16LL * num_sms());
  adagrad_table_kernel<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(p, state_sum, grad, row_touched, rows, clip, hv, hyper_dev);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_adagrad_dense(float* p, float* state_sum, const float* grad, int64_t n, const float* clip, const fbn_adagrad_t* h,
                                 const float* hyper_dev, fbn_stream_t stream) {
  FBN_REQUIRE(p && state_sum && grad && (h || hyper_dev), FBN_ERR_ARG, "fbn_adagrad_dense: null pointer");
  FBN_REQUIRE(n % 4 == 0, FBN_ERR_SHAPE, "fbn_adagrad_dense: n must be a multiple of 4 (pad the flat buffer)");
  FBN_REQUIRE(aligned16(p) && aligned16(state_sum) && aligned16(grad), FBN_ERR_ALIGN, "fbn_adagrad_dense: unaligned pointer");
  AdagradHyper hv{};
  if (h) hv = make_adagrad(*h);
  int blocks = (int)std::min<long long>(cdiv(n / 4, 256), 16LL * num_sms());
  adagrad_dense_kernel<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(p, state_sum, grad, n, clip, hv, hyper_dev);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

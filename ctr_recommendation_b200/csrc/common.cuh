// Shared device/host helpers for libfibinet_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fibinet_b200.h"

namespace fbn {

constexpr int D = FBN_D;        // 128 floats = 32 lanes x float4: one warp-wide 512 B row access
constexpr int NF = FBN_F;       // 6 fields; field 0 (user) is the constant zero vector (ref :152)
constexpr int NA = 5;           // active fields 1..5
constexpr int K1 = FBN_K1;      // 2688
constexpr int H1 = FBN_H1;
constexpr int H2 = FBN_H2;
constexpr int SE_R_DEFAULT = FBN_SE_R;   // the reference's SENetLayer(6, reduction_ratio=2)
constexpr int MAX_L = 64;
constexpr int MAX_CATE = 16;

void set_error(const char* fmt, ...);

#define FBN_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      fbn::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return FBN_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library (bench.py's gpu_launches claim)
#define FBN_CHECK_LAUNCH()               \
  do {                                   \
    ++fbn::g_launches;                   \
    FBN_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

#define FBN_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      fbn::set_error(__VA_ARGS__);        \
      return (code);                      \
    }                                     \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
int num_sms();

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float hsum4(const float4& a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4s(const float* p) {  // streaming read, do not pollute L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float a) { return make_float4(a, a, a, a); }
__device__ __forceinline__ float4 operator+(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 operator-(const float4& a, const float4& b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 operator*(const float4& a, const float4& b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 operator*(const float4& a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 operator/(const float4& a, float s) {
  return make_float4(a.x / s, a.y / s, a.z / s, a.w / s);
}
__device__ __forceinline__ void operator+=(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// store 4 consecutive values in tcgen05 operand format at element index `e` of a packed tensor (see PackDst in tower.h):
// mode FBN_PREC_TF32X3: hi = x & 0xffffe000 at e, lo = x - hi at e + lo_off ; FBN_PREC_BF16: bf16 round-to-nearest
__device__ __forceinline__ void store_packed4(void* base, long long lo_off, int mode, long long e, const float4& v) {
  if (mode == FBN_PREC_TF32X3) {
    float* d = static_cast<float*>(base);
    float4 h;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    st4(d + e, h);
    st4(d + lo_off + e, v - h);
  } else if (mode == FBN_PREC_BF16) {
    unsigned short* d = static_cast<unsigned short*>(base);
    auto rn = [](float x) {  // fp32 -> bf16 round-to-nearest-even (finite inputs)
      unsigned u = __float_as_uint(x);
      u += 0x7fffu + ((u >> 16) & 1u);
      return (unsigned)(u >> 16);
    };
    uint2 o;
    o.x = rn(v.x) | (rn(v.y) << 16);
    o.y = rn(v.z) | (rn(v.w) << 16);
    *reinterpret_cast<uint2*>(d + e) = o;
  }
}

// ---- FBN_PREC_F16X3 operand format (gemm_tc.cu) -----------------------------------------------------------------------------
// A packed tensor is followed by a small float record ("tail"): [0] = s (the tensor's power-of-two scale), [1] = 1 / s, [2] = amax,
// [4] = number of partial maxima, [F16_REC_FLOATS + i] = partial maximum i of |x| (written by an amax pass or, block by block, by
// the kernel that produces the tensor).  max is order-independent, so the scale is a deterministic function of the values alone.
constexpr int F16_REC_FLOATS = 16;
constexpr int F16_AMAX_BLOCKS = 1024;

__device__ __forceinline__ float f16x3_scale(float amax) {
  // amax * s in [2^14, 2^15) (fp16 overflows at 65504); an all-zero or non-finite tensor is left unscaled; the exponent is clamped so
  // that s and 1/s are normal fp32 numbers (only a tensor with amax < 2^-106 is scaled less than ideally)
  const unsigned bits = __float_as_uint(amax);
  const int e = (int)((bits >> 23) & 0xffu) - 127;
  if (bits == 0u || e == 128) return 1.f;
  const int k = max(-120, min(120, 14 - e));
  return __uint_as_float((unsigned)(k + 127) << 23);
}

// max over a 256-thread block (every thread must call); the result is valid in thread 0
__device__ __forceinline__ float block_max_256(float m) {
  __shared__ float sm_bmax[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm_bmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm_bmax[w]);
  }
  __syncthreads();      // the scratch may be reused by a second call
  return m;
}

// a producer kernel (256 threads per block, at most F16_AMAX_BLOCKS blocks) publishes its block's maximum of |values written|
__device__ __forceinline__ void f16x3_publish_amax(float m, float* tail) {
  m = block_max_256(m);
  if (threadIdx.x == 0) {
    tail[F16_REC_FLOATS + blockIdx.x] = m;
    if (blockIdx.x == 0) tail[4] = (float)gridDim.x;
  }
}

// every block of a pack kernel: fold the partial maxima into the tensor's scale (block 0 records it for the GEMM epilogues).
// npartial < 0: the count the producer left in tail[4].
__device__ __forceinline__ float f16x3_block_scale(float* tail, int npartial) {
  __shared__ float s_scale;
  if (npartial < 0) npartial = (int)tail[4];
  float m = 0.f;
  for (int i = threadIdx.x; i < npartial; i += 256) m = fmaxf(m, tail[F16_REC_FLOATS + i]);
  m = block_max_256(m);
  if (threadIdx.x == 0) {
    const float sc = f16x3_scale(m);
    s_scale = sc;
    if (blockIdx.x == 0) { tail[0] = sc; tail[1] = 1.0f / sc; tail[2] = m; }
  }
  __syncthreads();
  return s_scale;
}

// 4 consecutive values -> hi = fp16_rn(s x) at element e, lo = fp16_rn(s x - hi) at e + lo_off
__device__ __forceinline__ void store_f16x3_4(__half* dst, long long lo_off, long long e, float4 v, float sc) {
  v = v * sc;
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
  uint2 oh, ol;
  oh.x = *reinterpret_cast<const uint32_t*>(&h0); oh.y = *reinterpret_cast<const uint32_t*>(&h1);
  ol.x = *reinterpret_cast<const uint32_t*>(&l0); ol.y = *reinterpret_cast<const uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(dst + e) = oh;
  *reinterpret_cast<uint2*>(dst + lo_off + e) = ol;
}
__device__ __forceinline__ float amax4(const float4& v) { return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))); }

// tensor.long() on the reference's index columns (src/model_fibinet.py:140-143): truncation toward
// zero for floating inputs; exact for |id| < 2^53 (float64) as the loader delivers them.
__device__ __forceinline__ long long load_index(const void* p, int dtype, long long i) {
  switch (dtype) {
    case FBN_IDX_I32: return static_cast<const int32_t*>(p)[i];
    case FBN_IDX_I64: return static_cast<const long long*>(p)[i];
    case FBN_IDX_F64: return static_cast<long long>(static_cast<const double*>(p)[i]);
    default: return static_cast<long long>(static_cast<const float*>(p)[i]);
  }
}

// Philox4x32-10 counter-based generator: dropout stream keyed by (seed, offset), one call per
// 4 consecutive elements.  (The reference draws from torch's bernoulli_ stream, which cannot be
// reproduced; train-mode parity tests pass explicit keep-masks instead.)
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(ctr_lo), c1 = static_cast<uint32_t>(ctr_lo >> 32);
  uint32_t c2 = static_cast<uint32_t>(ctr_hi), c3 = static_cast<uint32_t>(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// Workspace layout (one contiguous scratch block per batch size, carved deterministically).
// ---------------------------------------------------------------------------------------------
struct Workspace {
  int64_t B, L;
  // canonical indices
  int32_t* ids;      // (B,4): item_id, likes, views, n_valid
  int32_t* seq;      // (B,L)
  int32_t* idflag;   // (4) sticky out-of-range flags written by the gather kernel
  // forward activations kept for backward
  float* X5;         // (B,5,128) fields 1..5 before SENET
  float* sgate;      // (B,8) sigmoid gates s_0..s_5
  float* xhat;       // (B,128) LayerNorm normalised projection
  float* xmm;        // (B,128) gathered item_emb_d128 rows (resident-table mode only)
  float* Ymm;        // (B,128) item_emb_d128 x mm_w^T + mm_b from the projection GEMM (batch-vector mode)
  float* rstd;       // (B)
  float* cnt;        // (B) history count clamp(min=1)
  float* C;          // (B,2688) MLP input [V | pairs]; blocks 0 and 6..10 stay zero
  float* T;          // (B,10,128) bilinear transforms (4 used for all/each)
  float* Hd1;        // (B,512) pre-BN
  float* A1;         // (B,512)
  float* Hd2;        // (B,256)
  float* A2;         // (B,256)
  float* logit;      // (B)
  float* prob;       // (B)
  float* bn;         // mean1[512] rstd1[512] mean2[256] rstd2[256]
  // backward
  float* dlogit;     // (B)
  float* dH2;        // (B,256)
  float* dH1;        // (B,512) (dA1 then dH1 in place)
  float* dC;         // (B,2688) active blocks only
  float* dT;         // (B,10,128)
  float* dV;         // (B,5,128)
  float* dXitem;     // (B,128)
  float* dXhist;     // (B,128) already divided by cnt
  float* dln;        // (B,128)
  float* dy;         // (B,128)
  float* sestat;     // (B,24)
  float* partial;    // reduction partials (main stream)
  size_t partial_floats;
  float* partial_side;   // same size: scratch of the leaf computations that run on the library's side stream
  float* partial_cate;   // per-CTA cate_emb gradient partials (produced on the main stream, reduced on the side stream)
  float* partial_embsq;  // per-CTA sum-of-squares partials of the table gradient
  // embedding backward
  int32_t* keys_in;  // (B*(L+1))
  int32_t* keys_out;
  int32_t* vals_in;
  int32_t* vals_out;
  int32_t* row_off;  // (item_rows + 1)
  int32_t* row_cnt;  // (item_rows + 1)
  void* cub_tmp;
  size_t cub_bytes;
  void* gemm_scratch;   // operand scratch for GEMMs whose inputs are not pre-packed
  size_t gemm_scratch_bytes;
  // tcgen05 operands packed once per step (tf32 hi|lo or bf16), natural layout
  void* pk_C; void* pk_A1; void* pk_dH2; void* pk_dH1; void* pk_dT; void* pk_dy; void* pk_xmm;
  void* pk_w1; void* pk_w2; void* pk_bil; void* pk_mmw;
  void* pk16_C;      // f16x3: the live MLP-input row as fp16 hi|lo (pk_C then holds the tf32 copy of the field blocks only)
  size_t total_bytes;
};

// Fills w from (base, B, L, item_rows). base may be nullptr to just compute total_bytes.
void carve_workspace(Workspace& w, void* base, int64_t B, int64_t L, int64_t rows);

}  // namespace fbn

// Internal GEMM descriptor shared by the SIMT (fp32) and tcgen05 (tf32x3 / bf16) back ends.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fbn {

// An operand already converted to the tcgen05 format (tf32 hi|lo split, fp16 hi|lo split under one scale, or bf16) in its natural
// row-major layout.
struct Packed {
  void* data = nullptr;      // hi part (or the bf16 copy); (rows, pitch) elements
  long long pitch = 0;       // elements per row (multiple of 8)
  long long lo_off = 0;      // element offset of the lo part (tf32x3)
  long long rows = 0, cols = 0;
  float* scale = nullptr;    // f16x3: device record {s, 1/s, amax} of the tensor's power-of-two scale (written by the pack pass)
  Packed view_cols(long long col0, int esz) const {   // column-block view (same pitch / lo_off)
    Packed p = *this;
    p.data = static_cast<char*>(data) + col0 * esz;
    return p;
  }
};

// C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N])
//   a_t == 0: A stored (M,K) row-major, lda ; a_t != 0: A stored (K,M) row-major, lda
//   b_t == 0: B stored (K,N) row-major, ldb ; b_t != 0: B stored (N,K) row-major, ldb
// blockIdx.z enumerates batch x splits: batch index advances A/B/C by stride{A,B,C}; split s covers
// a contiguous range of K tiles and writes to C + s*strideSplit (reduce afterwards).
struct GemmArgs {
  const float* A = nullptr;
  const float* B = nullptr;
  const float* bias = nullptr;
  float* C = nullptr;
  long long M = 0, N = 0, K = 0;
  long long lda = 0, ldb = 0, ldc = 0;
  int a_t = 0, b_t = 0;
  int batch = 1, splits = 1;
  long long strideA = 0, strideB = 0, strideC = 0, strideSplit = 0;
  unsigned long long kmask = ~0ull;  // bit i: 128-wide K block i is non-zero
  unsigned long long nmask = ~0ull;  // bit i: 128-wide N block i is needed
  int accumulate = 0;
  // optional pre-packed operands (tcgen05 precisions): same storage orientation as A / B; batch index advances them by
  // strideA / strideB elements like the fp32 pointers
  Packed pkA, pkB;
};

int gemm_simt(const GemmArgs& g, cudaStream_t st);
// tcgen05 back end; returns FBN_ERR_SHAPE if the shape/layout is not covered (caller falls back is NOT
// allowed silently: the dispatcher reports the error).
int gemm_tc(const GemmArgs& g, int precision, void* scratch, size_t scratch_bytes, cudaStream_t st);
bool gemm_tc_supported(const GemmArgs& g, int precision);
size_t gemm_tc_scratch_bytes(long long M, long long N, long long K, int precision);
size_t packed_bytes(long long rows, long long cols, int precision);
Packed packed_describe(void* region, long long rows, long long cols, int precision = -1);
// producer_amax (f16x3 only): the kernel that wrote `src` already published its partial maxima into the region's record
// (PackDst::tail), so the amax pass is skipped
int pack_operand(const float* src, long long ld, long long rows, long long cols, int precision, void* dst, unsigned long long colmask,
                 Packed* out, cudaStream_t st, bool producer_amax = false);
int gemm(const GemmArgs& g, int precision, void* scratch, size_t scratch_bytes, cudaStream_t st);

}  // namespace fbn

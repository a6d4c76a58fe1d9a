// E1: deterministic sorted-segment embedding backward (reference: embedding_dense_backward of the
// four nn.Embedding lookups on item_emb, src/model_fibinet.py:159,167, padding_idx=0 at :100).
//
// Occurrences (B target ids + B*L history ids) are keyed by table row, stably radix-sorted so that
// every row's occurrences are in source order, and each table row is then summed by ONE warp in that
// fixed order: no floating point atomics, bitwise reproducible run to run.  Padding id 0 is mapped to
// a sentinel key beyond the last row and never summed (its gradient is exactly zero, like torch).
// The output is the dense (rows,128) gradient the reference's dense Adam consumes.
#include <cub/cub.cuh>

#include "common.cuh"
#include "tower.h"

namespace fbn {

__global__ void emb_build_keys_kernel(const void* __restrict__ item_id, int idx_dtype, const void* __restrict__ seq, int seq_dtype,
                                      long long B, int L, long long rows, int32_t* __restrict__ keys, int32_t* __restrict__ vals,
                                      int32_t* __restrict__ row_count) {
  const long long n = B * (1 + (seq ? L : 0));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long raw = i < B ? load_index(item_id, idx_dtype, i) : load_index(seq, seq_dtype, i - B);
    int key = (int)min(max(raw, 0LL), rows - 1);   // same clamp as the forward gather
    if (key <= 0) key = (int)rows;                 // padding -> sentinel beyond the last row
    else atomicAdd(row_count + key, 1);            // integer atomics: order-independent result
    keys[i] = key;
    vals[i] = (int)i;
  }
}

constexpr int ER_WARPS = 8;

__global__ void __launch_bounds__(ER_WARPS * 32) emb_rows_kernel(const int32_t* __restrict__ row_count,
                                                                 const int32_t* __restrict__ row_off,
                                                                 const int32_t* __restrict__ src, const float* __restrict__ dXitem,
                                                                 const float* __restrict__ dXhist, long long B, int L, long long rows,
                                                                 int zero_fill, float* __restrict__ grad,
                                                                 float* __restrict__ sumsq_partial) {
  __shared__ float s_sq[ER_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * ER_WARPS + warp;
  float sq = 0.f;
  if (r < rows) {
    const int cnt = __ldg(row_count + r), off = __ldg(row_off + r);   // independent loads, one latency
    if (cnt > 0) {
      float4 acc = f4(0.f);
      for (int o0 = 0; o0 < cnt; o0 += 32) {
        const int mine = (o0 + lane < cnt) ? __ldg(src + off + o0 + lane) : 0;
        const int n = min(32, cnt - o0);
        for (int k = 0; k < n; k += 4) {            // 4 independent 512-byte row loads in flight, summed in source order
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int s = __shfl_sync(0xffffffffu, mine, min(k + u, 31));
            const float* p = s < B ? dXitem + (long long)s * D : dXhist + ((long long)(s - B) / L) * D;
            v[u] = (k + u < n) ? ld4(p + 4 * lane) : f4(0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < n) acc += v[u];
        }
      }
      st4(grad + r * D + 4 * lane, acc);
      sq = warp_sum(hsum4(acc * acc));
    } else if (zero_fill) {
      st4(grad + r * D + 4 * lane, f4(0.f));
    }
  }
  if (lane == 0) s_sq[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ER_WARPS; ++w) t += s_sq[w];
    sumsq_partial[blockIdx.x] = t;
  }
}

__global__ void sum_partials_kernel(const float* __restrict__ x, long long n, float* out) {
  __shared__ double s[256];
  double t = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) t += (double)x[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)s[0];
}

static int key_bits(long long rows) {
  int b = 1;
  while ((1LL << b) <= rows) ++b;
  return b;
}

size_t emb_sort_temp_bytes(long long n, long long rows) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (int)std::max<long long>(n, 1), 0, key_bits(rows));
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (int)std::max<long long>(rows, 1));
  return std::max(a, b) + 256;
}

int emb_grad_partial_count(long long rows) { return (int)cdiv(rows, ER_WARPS); }

// stage 1 (depends on the batch ids only -- can run concurrently with the forward pass): occurrences keyed by table row,
// stably sorted, per-row counts and offsets
int emb_index(const EmbGradArgs& a, cudaStream_t st) {
  const long long n = a.B * (1 + (a.seq ? a.L : 0));
  FBN_REQUIRE(n < (1LL << 31) && a.rows < (1LL << 30), FBN_ERR_SHAPE, "embedding backward: too many occurrences");
  FBN_CHECK_CUDA(cudaMemsetAsync(a.row_count, 0, sizeof(int32_t) * a.rows, st));
  int blocks = (int)std::min<long long>(cdiv(n, 256), 8LL * num_sms());
  emb_build_keys_kernel<<<std::max(blocks, 1), 256, 0, st>>>(a.item_id, a.idx_dtype, a.seq, a.seq_dtype, a.B, a.L, a.rows, a.keys_in,
                                                             a.vals_in, a.row_count);
  FBN_CHECK_LAUNCH();
  size_t bytes = a.cub_bytes;
  FBN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(a.cub_tmp, bytes, (const int32_t*)a.keys_in, a.keys_out, (const int32_t*)a.vals_in,
                                                 a.vals_out, (int)n, 0, key_bits(a.rows), st));
  g_launches += 4;  // cub: histogram + exclusive-sum + onesweep passes (17-bit keys)
  bytes = a.cub_bytes;
  FBN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(a.cub_tmp, bytes, (const int32_t*)a.row_count, a.row_off, (int)a.rows, st));
  g_launches += 2;  // cub: scan init + scan
  return FBN_OK;
}

// stage 2: one warp per table row sums its occurrences in source order
int emb_rows(const EmbGradArgs& a, cudaStream_t st) {
  const int nb = emb_grad_partial_count(a.rows);
  emb_rows_kernel<<<nb, ER_WARPS * 32, 0, st>>>(a.row_count, a.row_off, a.vals_out, a.dXitem, a.dXhist, a.B, a.L > 0 ? a.L : 1, a.rows,
                                                a.zero_fill, a.grad, a.sumsq_partial);
  FBN_CHECK_LAUNCH();
  sum_partials_kernel<<<1, 256, 0, st>>>(a.sumsq_partial, nb, a.sumsq_out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

}  // namespace fbn

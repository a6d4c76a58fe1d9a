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
#include "segsum.cuh"
#include "tower.h"

namespace fbn {

__global__ void emb_build_keys_kernel(const void* __restrict__ item_id, int idx_dtype, const void* __restrict__ seq, int seq_dtype,
                                      long long B, int L, long long rows, int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const long long n = B * (1 + (seq ? L : 0));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long raw = i < B ? load_index(item_id, idx_dtype, i) : load_index(seq, seq_dtype, i - B);
    int key = (int)min(max(raw, 0LL), rows - 1);   // same clamp as the forward gather
    if (key <= 0) key = (int)rows;                 // padding -> sentinel beyond the last row
    keys[i] = key;
    vals[i] = (int)i;
  }
}

// per-row occurrence count and offset from the SORTED keys: the thread at the last element of a run finds the run's start by
// binary search (no atomics: a Zipf head row owning 8 % of the batch would serialise tens of thousands of them)
__global__ void emb_runs_kernel(const int32_t* __restrict__ keys, long long n, int rows, int32_t* __restrict__ row_count,
                                int32_t* __restrict__ row_off) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = keys[i];
    if (k >= rows || (i + 1 < n && keys[i + 1] == k)) continue;
    long long lo = 0, hi = i;     // first position holding k
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    row_off[k] = (int)lo;
    row_count[k] = (int)(i + 1 - lo);
  }
}

static int key_bits(long long rows) {
  int b = 1;
  while ((1LL << b) <= rows) ++b;
  return b;
}

// scratch of the embedding backward: radix-sort temporaries followed by the hot-row work lists of segsum.cuh
static size_t sort_bytes(long long n, long long rows) {
  size_t a = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (int)std::max<long long>(n, 1), 0, key_bits(rows));
  return (a + 511) & ~size_t(255);
}
size_t emb_sort_temp_bytes(long long n, long long rows) { return sort_bytes(n, rows) + seg_scratch_bytes(n); }

int emb_grad_partial_count(long long rows) { return (int)std::min<long long>(cdiv(rows, SEG_WARPS), 16LL * 148); }

// stage 1 (depends on the batch ids only -- can run concurrently with the forward pass): occurrences keyed by table row,
// stably sorted, per-row counts and offsets
int emb_index(const EmbGradArgs& a, cudaStream_t st) {
  const long long n = a.B * (1 + (a.seq ? a.L : 0));
  FBN_REQUIRE(n < (1LL << 31) && a.rows < (1LL << 30), FBN_ERR_SHAPE, "embedding backward: too many occurrences");
  FBN_CHECK_CUDA(cudaMemsetAsync(a.row_count, 0, sizeof(int32_t) * a.rows, st));
  int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n, 256), 8LL * num_sms()));
  emb_build_keys_kernel<<<blocks, 256, 0, st>>>(a.item_id, a.idx_dtype, a.seq, a.seq_dtype, a.B, a.L, a.rows, a.keys_in, a.vals_in);
  FBN_CHECK_LAUNCH();
  size_t bytes = sort_bytes(n, a.rows);
  FBN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(a.cub_tmp, bytes, (const int32_t*)a.keys_in, a.keys_out, (const int32_t*)a.vals_in,
                                                 a.vals_out, (int)n, 0, key_bits(a.rows), st));
  g_launches += 4;  // cub: histogram + exclusive-sum + onesweep passes (17-bit keys)
  emb_runs_kernel<<<blocks, 256, 0, st>>>(a.keys_out, n, (int)a.rows, a.row_count, a.row_off);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// stage 2: one warp per table row sums its occurrences in source order (hot rows: chunked, segsum.cuh)
int emb_rows(const EmbGradArgs& a, cudaStream_t st) {
  const long long n = a.B * (1 + (a.seq ? a.L : 0));
  const int nb = emb_grad_partial_count(a.rows);
  SegArgs s{};
  s.off = a.row_off; s.cnt = a.row_count; s.nseg_dev = nullptr; s.nseg = a.rows; s.src = a.vals_out;
  s.dXitem = a.dXitem; s.dXhist = a.dXhist; s.B = a.B; s.L = a.L > 0 ? a.L : 1;
  s.out = a.grad; s.zero_fill = a.zero_fill; s.sq_partial = a.sumsq_partial; s.nseg_bound = a.rows;
  s.hot = seg_carve(static_cast<char*>(a.cub_tmp) + sort_bytes(n, a.rows), n);
  FBN_CHECK_CUDA(seg_sum_launch(s, nb, n, st));
  g_launches += 3;
  seg_sumsq_final_kernel<<<1, 256, 0, st>>>(a.sumsq_partial, nb, s.hot.hot_sq, (int)s.hot.max_hot, a.sumsq_out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

}  // namespace fbn

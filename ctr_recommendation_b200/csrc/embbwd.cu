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

// ---- F-field model (general.py): lookups into tables stored back to back, and their dense gradient -----------------------------
// Field descriptor (5 x int64 per field, on the device): {first row of its table, vocabulary, first id column, bag length, padding id}.
//   bag length 1 : x[b][f] = table[row0 + id]                                   (nn.Embedding lookup, ref :155-159)
//   bag length L : x[b][f] = sum over the L ids != padding of their rows / max(count, 1)   (the reference's masked mean pooling of
//                  item_seq, ref :165-174; item_tags is a bag of 5)
// Several fields may name the same table (ref: likes_level / views_level share cate_emb, item_id / item_seq share item_emb).
// An id equal to the padding id contributes nothing and receives no gradient (nn.Embedding(padding_idx=...), ref :100); -1 = none.
constexpr int FDESC = 5;

__global__ void fields_gather_kernel(const float* __restrict__ table, const long long* __restrict__ desc, const void* __restrict__ ids,
                                     int idx_dtype, long long B, int F, int cols, float* __restrict__ x, float* __restrict__ cnt,
                                     int32_t* flag) {
  const int lane = threadIdx.x & 31;
  const long long n = B * F;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {   // one warp per (sample, field)
    const int f = (int)(i % F);
    const long long b = i / F;
    const long long* d = desc + f * FDESC;
    const long long row0 = d[0], vocab = d[1], c0 = d[2], pad = d[4];
    const int L = (int)d[3];
    float4 acc = f4(0.f);
    int nvalid = 0;
    for (int l0 = 0; l0 < L; l0 += 8) {               // 8 independent row loads in flight, accumulated in id order
      long long id[8];
      float4 r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        id[u] = pad;
        if (l0 + u < L) {
          long long v = load_index(ids, idx_dtype, b * cols + c0 + l0 + u);
          if (lane == 0 && (v < 0 || v >= vocab)) flag[0] = 1;        // torch: IndexError (the host raises it)
          id[u] = min(max(v, 0LL), vocab - 1);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = (l0 + u < L && id[u] != pad) ? ld4(table + (row0 + id[u]) * D + 4 * lane) : f4(0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (l0 + u < L && id[u] != pad) { acc += r[u]; ++nvalid; }
    }
    const float c = (float)max(nvalid, 1);
    st4(x + i * D + 4 * lane, L == 1 ? acc : acc / c);
    if (lane == 0) cnt[i] = L == 1 ? 1.f : c;
  }
}

// occurrence keys of the dense gradient: every id column of every sample; padding ids go to the sentinel key (never summed)
__global__ void fields_build_keys_kernel(const long long* __restrict__ desc, const int32_t* __restrict__ colfield, const void* __restrict__ ids,
                                         int idx_dtype, long long n, int F, int cols, long long rows, int32_t* __restrict__ keys,
                                         int32_t* __restrict__ vals) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % cols);
    const long long b = i / cols;
    const int f = colfield[col];
    const long long* d = desc + f * FDESC;
    const long long id = min(max(load_index(ids, idx_dtype, i), 0LL), d[1] - 1);
    keys[i] = id == d[4] ? (int)rows : (int)(d[0] + id);
    vals[i] = (int)(b * F + f);                      // the gradient row this occurrence adds: dx[b][f] (already divided by the bag count)
  }
}

__global__ void fields_scale_kernel(float* __restrict__ dx, const float* __restrict__ cnt, long long n) {   // dx[b][f] /= count[b][f]
  const int lane = threadIdx.x & 31;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
    const float c = __ldg(cnt + i);
    if (c != 1.f) st4(dx + i * D + 4 * lane, ld4s(dx + i * D + 4 * lane) / c);
  }
}

struct FieldsScratch { int32_t *keys_in, *keys_out, *vals_in, *vals_out, *row_off, *row_cnt; float* sq_partial; void* cub; size_t total; };

static FieldsScratch fields_carve(void* base, long long n, long long rows) {
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  FieldsScratch s;
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
  char* p0 = p;
  s.keys_in = (int32_t*)p; p += al(n * 4);
  s.keys_out = (int32_t*)p; p += al(n * 4);
  s.vals_in = (int32_t*)p; p += al(n * 4);
  s.vals_out = (int32_t*)p; p += al(n * 4);
  s.row_off = (int32_t*)p; p += al((rows + 1) * 4);
  s.row_cnt = (int32_t*)p; p += al((rows + 1) * 4);
  s.sq_partial = (float*)p; p += al(((size_t)cdiv(rows, 8) + 1024) * 4);
  s.cub = p; p += emb_sort_temp_bytes(n, rows);
  s.total = (size_t)(p - p0) + 256;
  return s;
}

}  // namespace fbn

using namespace fbn;

extern "C" int fbn_fields_gather(const float* table, const int64_t* desc, const void* ids, int idx_dtype, int64_t batch, int fields,
                                 int id_cols, float* x, float* count, int32_t* flag, fbn_stream_t stream) {
  FBN_REQUIRE(table && desc && ids && x && count && flag, FBN_ERR_ARG, "fbn_fields_gather: null pointer");
  FBN_REQUIRE(idx_dtype == FBN_IDX_I32 || idx_dtype == FBN_IDX_I64, FBN_ERR_DTYPE, "fbn_fields_gather: ids must be int32 or int64");
  FBN_REQUIRE(fields >= 1 && fields <= 64 && batch >= 1 && id_cols >= fields, FBN_ERR_SHAPE, "fbn_fields_gather: need 1 <= fields <= 64");
  FBN_REQUIRE(aligned16(table) && aligned16(x), FBN_ERR_ALIGN, "fbn_fields_gather: unaligned pointer");
  const long long n = (long long)batch * fields;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n, 8), 16LL * num_sms()));
  fields_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(table, reinterpret_cast<const long long*>(desc), ids, idx_dtype, batch, fields,
                                                                 id_cols, x, count, flag);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" size_t fbn_fields_scatter_bytes(int64_t batch, int id_cols, int64_t rows) {
  return fields_carve(nullptr, (long long)batch * id_cols, rows).total;
}

extern "C" int fbn_fields_scatter(float* dx, const float* count, const int64_t* desc, const int32_t* col_field, const void* ids, int idx_dtype,
                                  int64_t batch, int fields, int id_cols, int64_t rows, float* grad, int32_t* row_touched, int zero_fill,
                                  float* sumsq_out, void* scratch, size_t scratch_bytes, fbn_stream_t stream) {
  FBN_REQUIRE(dx && count && desc && col_field && ids && grad && sumsq_out && scratch, FBN_ERR_ARG, "fbn_fields_scatter: null pointer");
  FBN_REQUIRE(idx_dtype == FBN_IDX_I32 || idx_dtype == FBN_IDX_I64, FBN_ERR_DTYPE, "fbn_fields_scatter: ids must be int32 or int64");
  const long long n = (long long)batch * id_cols;
  FBN_REQUIRE(fields >= 1 && fields <= 64 && id_cols >= fields && batch >= 1 && rows >= 1 && n < (1LL << 31) && rows < (1LL << 30) &&
                  (long long)batch * fields < (1LL << 31), FBN_ERR_SHAPE, "fbn_fields_scatter: bad shape");
  FBN_REQUIRE(aligned16(dx) && aligned16(grad), FBN_ERR_ALIGN, "fbn_fields_scatter: unaligned pointer");
  FBN_REQUIRE(scratch_bytes >= fbn_fields_scatter_bytes(batch, id_cols, rows), FBN_ERR_ARG, "fbn_fields_scatter: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  FieldsScratch s = fields_carve(scratch, n, rows);
  int32_t* cnt = row_touched ? row_touched : s.row_cnt;
  FBN_CHECK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * rows, st));
  {
    const long long nf = (long long)batch * fields;
    const int sb = (int)std::max<long long>(1, std::min<long long>(cdiv(nf, 8), 8LL * num_sms()));
    fields_scale_kernel<<<sb, 256, 0, st>>>(dx, count, nf);          // mean pooling: every id of a bag receives dx / count (ref :174)
    FBN_CHECK_LAUNCH();
  }
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n, 256), 8LL * num_sms()));
  fields_build_keys_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(desc), col_field, ids, idx_dtype, n, fields, id_cols, rows,
                                                   s.keys_in, s.vals_in);
  FBN_CHECK_LAUNCH();
  size_t bytes = sort_bytes(n, rows);
  FBN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(s.cub, bytes, (const int32_t*)s.keys_in, s.keys_out, (const int32_t*)s.vals_in, s.vals_out,
                                                 (int)n, 0, key_bits(rows), st));
  g_launches += 4;
  emb_runs_kernel<<<blocks, 256, 0, st>>>(s.keys_out, n, (int)rows, cnt, s.row_off);
  FBN_CHECK_LAUNCH();
  // occurrence value s = gradient row index b * F + f: B = 0 and L = 1 in the shared segment-sum addressing reads dx[s]
  const int nb = emb_grad_partial_count(rows);
  SegArgs a{};
  a.off = s.row_off; a.cnt = cnt; a.nseg_dev = nullptr; a.nseg = rows; a.src = s.vals_out;
  a.dXitem = dx; a.dXhist = dx; a.B = 0; a.L = 1;
  a.out = grad; a.zero_fill = zero_fill; a.sq_partial = s.sq_partial; a.nseg_bound = rows;
  a.hot = seg_carve(static_cast<char*>(s.cub) + sort_bytes(n, rows), n);
  FBN_CHECK_CUDA(seg_sum_launch(a, nb, n, st));
  g_launches += 3;
  seg_sumsq_final_kernel<<<1, 256, 0, st>>>(s.sq_partial, nb, a.hot.hot_sq, (int)a.hot.max_hot, sumsq_out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// Deterministic sorted-segment sum of embedding gradient rows, shared by the replicated table (embbwd.cu) and the row-sharded
// table (shard.cu).  Reference semantics: embedding_dense_backward of the item_emb lookups (src/model_fibinet.py:159,167).
//
// Input: occurrences sorted by output row; segment i = src[off(i) .. off(i) + cnt(i)) lists, in source order, the samples
// whose gradient row (dXitem[s] for s < B, dXhist[(s - B) / L] otherwise) is added to output row i.
//
// The value of a row depends only on its own occurrence list, never on the grid, the block boundaries or which rows happen
// to be neighbours -- that is what makes replicas, re-runs and the replicated / sharded paths agree bit for bit:
//   cnt <= SEG_HOT : one warp adds the rows sequentially in source order;
//   cnt >  SEG_HOT : ("hot" row, Zipf heads: one id can own 8 % of a batch) the list is cut into chunks of SEG_HOT
//                    occurrences *relative to the row's own start*, every chunk is summed sequentially by its own warp
//                    anywhere on the GPU, and one warp then adds the chunk sums in chunk order.
// Hot rows are discovered by the first kernel and appended to a work list with atomics; the order of that list is arbitrary
// but no value depends on it (the sum-of-squares contributions are slotted by the rank of the row id).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace fbn {

constexpr int SEG_HOT = 256;
constexpr int SEG_WARPS = 8;

struct SegScratch {
  int32_t* ctr;        // [0] = number of hot rows, [1] = number of chunk entries   (zeroed every step)
  int32_t* hot_seg;    // (max_hot) segment index
  int32_t* hot_base;   // (max_hot) first chunk entry
  int32_t* chunk_seg;  // (max_chunks) segment index of the chunk
  int32_t* chunk_idx;  // (max_chunks) chunk number inside its segment
  float* chunk_sum;    // (max_chunks, 128)
  float* hot_sq;       // (max_hot) sum of squares of hot rows, slotted by rank of the segment index (zeroed every step)
  long long max_hot, max_chunks;
};

inline long long seg_max_hot(long long n) { return n / SEG_HOT + 1; }
inline long long seg_max_chunks(long long n) { return n / SEG_HOT + seg_max_hot(n) + 1; }
inline size_t seg_scratch_bytes(long long n) {
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  const long long h = seg_max_hot(n), c = seg_max_chunks(n);
  return al(16) + 2 * al(h * 4) + 2 * al(c * 4) + al((size_t)c * D * 4) + al(h * 4);
}
inline SegScratch seg_carve(void* base, long long n) {
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  SegScratch s;
  s.max_hot = seg_max_hot(n); s.max_chunks = seg_max_chunks(n);
  char* p = static_cast<char*>(base);
  s.ctr = (int32_t*)p; p += al(16);
  s.hot_sq = (float*)p; p += al(s.max_hot * 4);          // directly after ctr: one memset clears both
  s.hot_seg = (int32_t*)p; p += al(s.max_hot * 4);
  s.hot_base = (int32_t*)p; p += al(s.max_hot * 4);
  s.chunk_seg = (int32_t*)p; p += al(s.max_chunks * 4);
  s.chunk_idx = (int32_t*)p; p += al(s.max_chunks * 4);
  s.chunk_sum = (float*)p;
  return s;
}
inline size_t seg_clear_bytes(long long n) {   // ctr + hot_sq
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  return al(16) + al(seg_max_hot(n) * 4);
}

struct SegArgs {
  const int32_t* off;      // segment starts
  const int32_t* cnt;      // segment lengths, or nullptr: cnt(i) = off[i + 1] - off[i]
  const int32_t* nseg_dev; // number of segments on the device, or nullptr: nseg
  long long nseg;
  const int32_t* src;      // sorted occurrence ids
  const float* dXitem; const float* dXhist; long long B; int L;
  float* out;              // (nseg, 128)
  int zero_fill;           // write zeros to empty segments
  float* sq_partial;       // per-CTA sums of squares of the first kernel (gridDim.x entries)
  long long nseg_bound;    // host-side upper bound of the number of non-empty segments (picks the segments-per-warp variant)
  SegScratch hot;
};

#ifdef __CUDACC__
// sum of occurrences [o_begin, o_end) of one segment, sequentially in source order, 4 row loads in flight
__device__ __forceinline__ float4 seg_sum_range(const SegArgs& a, int o_begin, int o_end, int lane) {
  float4 acc = f4(0.f);
  for (int o0 = o_begin; o0 < o_end; o0 += 32) {
    const int mine = (o0 + lane < o_end) ? __ldg(a.src + o0 + lane) : 0;
    const int n = min(32, o_end - o0);
    for (int k = 0; k < n; k += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s = __shfl_sync(0xffffffffu, mine, min(k + u, 31));
        const float* p = s < a.B ? a.dXitem + (long long)s * D : a.dXhist + ((long long)(s - a.B) / a.L) * D;
        v[u] = (k + u < n) ? ld4(p + 4 * lane) : f4(0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k + u < n) acc += v[u];
    }
  }
  return acc;
}

__device__ __forceinline__ int seg_cnt(const SegArgs& a, long long i) {
  return a.cnt ? __ldg(a.cnt + i) : __ldg(a.off + i + 1) - __ldg(a.off + i);
}

// kernel 1 (default): one warp per segment; hot segments are only registered
static __global__ void __launch_bounds__(SEG_WARPS * 32) seg_rows_simple_kernel(SegArgs a) {
  __shared__ float s_sq[SEG_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nseg = a.nseg_dev ? (long long)*a.nseg_dev : a.nseg;
  float sq = 0.f;
  for (long long i = (long long)blockIdx.x * SEG_WARPS + warp; i < nseg; i += (long long)gridDim.x * SEG_WARPS) {
    const int cnt = seg_cnt(a, i);
    if (cnt > SEG_HOT) {
      const int nch = (cnt + SEG_HOT - 1) / SEG_HOT;
      int slot = 0, base = 0;
      if (lane == 0) {
        slot = atomicAdd(a.hot.ctr + 0, 1);
        base = atomicAdd(a.hot.ctr + 1, nch);
        a.hot.hot_seg[slot] = (int)i;
        a.hot.hot_base[slot] = base;
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int c = lane; c < nch; c += 32) {
        a.hot.chunk_seg[base + c] = (int)i;
        a.hot.chunk_idx[base + c] = c;
      }
    } else if (cnt > 0) {
      const int off = __ldg(a.off + i);
      const float4 acc = seg_sum_range(a, off, off + cnt, lane);
      st4(a.out + i * D + 4 * lane, acc);
      sq += warp_sum(hsum4(acc * acc));
    } else if (a.zero_fill) {
      st4(a.out + i * D + 4 * lane, f4(0.f));
    }
  }
  if (lane == 0) s_sq[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < SEG_WARPS; ++w) t += s_sq[w];
    a.sq_partial[blockIdx.x] = t;
  }
}

// kernel 1 (compact lists of short segments: a row-sharded 100 M-row table sees almost every id once): one warp per block of
// SEG_BLK consecutive segments.  Measured on B200 (one GPU, 20 M-row table, B = 65536): 1.35 M segments in 269 us where the per-segment kernel needed 358 us for 1.05 M; NOT used for the replicated
// table (B = 65536, 15 occurrences per row: 168 us vs 101 us for the per-segment kernel above -- the cursor bookkeeping costs more
// than the extra loads in flight buy -- and eight 200-occurrence Zipf segments in one warp would serialise 1600 row loads).  The occurrences of the block's ordinary segments are walked as one
// stream, 8 at a time: 8 source ids, then 8 independent 512-byte row loads, then the adds in stream order with a store at every
// segment end -- so a table with a million single-occurrence rows (row-sharded tables) keeps 8 rows in flight per warp instead
// of paying four dependent memory latencies per row, while the order of the adds inside a segment is unchanged.
// Hot segments are only registered here.
constexpr int SEG_ILP = 8;

template <int SEG_BLK>
static __global__ void __launch_bounds__(SEG_WARPS * 32) seg_rows_kernel(SegArgs a) {
  __shared__ float s_sq[SEG_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nseg = a.nseg_dev ? (long long)*a.nseg_dev : a.nseg;
  const long long nblk = (nseg + SEG_BLK - 1) / SEG_BLK;
  float sq = 0.f;
  for (long long blk = (long long)blockIdx.x * SEG_WARPS + warp; blk < nblk; blk += (long long)gridDim.x * SEG_WARPS) {
    const long long i0 = blk * SEG_BLK;
    const bool valid = lane < SEG_BLK && i0 + lane < nseg;
    const int my_cnt = valid ? seg_cnt(a, i0 + lane) : 0;
    const int my_off = my_cnt > 0 ? __ldg(a.off + i0 + lane) : 0;
    // empty segments / hot segments
    if (a.zero_fill) {
      unsigned empty = __ballot_sync(0xffffffffu, valid && my_cnt == 0);
      while (empty) {
        const int s = __ffs(empty) - 1;
        empty &= empty - 1;
        st4(a.out + (i0 + s) * D + 4 * lane, f4(0.f));
      }
    }
    unsigned hot = __ballot_sync(0xffffffffu, my_cnt > SEG_HOT);
    while (hot) {
      const int s = __ffs(hot) - 1;
      hot &= hot - 1;
      const int nch = (__shfl_sync(0xffffffffu, my_cnt, s) + SEG_HOT - 1) / SEG_HOT;
      int base = 0;
      if (lane == 0) {
        const int slot = atomicAdd(a.hot.ctr + 0, 1);
        base = atomicAdd(a.hot.ctr + 1, nch);
        a.hot.hot_seg[slot] = (int)(i0 + s);
        a.hot.hot_base[slot] = base;
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int c = lane; c < nch; c += 32) {
        a.hot.chunk_seg[base + c] = (int)(i0 + s);
        a.hot.chunk_idx[base + c] = c;
      }
    }
    // ordinary segments as one occurrence stream
    int s = 0, pos = 0;           // warp-uniform cursor: segment in the block, occurrence inside it
    float4 acc = f4(0.f);
    while (true) {
      int o[SEG_ILP], seg_of[SEG_ILP];
      bool last[SEG_ILP];
      int n = 0;
#pragma unroll
      for (int u = 0; u < SEG_ILP; ++u) {
        o[u] = 0; seg_of[u] = 0; last[u] = false;
        while (s < SEG_BLK) {
          const int c = __shfl_sync(0xffffffffu, my_cnt, s);
          if (c > 0 && c <= SEG_HOT && pos < c) break;
          ++s; pos = 0;
        }
        if (s < SEG_BLK) {
          const int c = __shfl_sync(0xffffffffu, my_cnt, s);
          o[u] = __shfl_sync(0xffffffffu, my_off, s) + pos;
          seg_of[u] = s;
          last[u] = pos == c - 1;
          ++pos;
          n = u + 1;
        }
      }
      if (n == 0) break;
      int id[SEG_ILP];
#pragma unroll
      for (int u = 0; u < SEG_ILP; ++u) id[u] = u < n ? __ldg(a.src + o[u]) : 0;
      float4 v[SEG_ILP];
#pragma unroll
      for (int u = 0; u < SEG_ILP; ++u) {
        const float* p = id[u] < a.B ? a.dXitem + (long long)id[u] * D : a.dXhist + ((long long)(id[u] - a.B) / a.L) * D;
        v[u] = u < n ? ld4(p + 4 * lane) : f4(0.f);
      }
#pragma unroll
      for (int u = 0; u < SEG_ILP; ++u) {
        if (u < n) {
          acc += v[u];
          if (last[u]) {
            st4(a.out + (i0 + seg_of[u]) * D + 4 * lane, acc);
            sq += warp_sum(hsum4(acc * acc));
            acc = f4(0.f);
          }
        }
      }
    }
  }
  if (lane == 0) s_sq[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < SEG_WARPS; ++w) t += s_sq[w];
    a.sq_partial[blockIdx.x] = t;
  }
}

// kernel 2: one warp per chunk of a hot segment
static __global__ void __launch_bounds__(SEG_WARPS * 32) seg_hot_chunks_kernel(SegArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int total = a.hot.ctr[1];
  for (int e = blockIdx.x * SEG_WARPS + warp; e < total; e += gridDim.x * SEG_WARPS) {
    const int i = a.hot.chunk_seg[e], c = a.hot.chunk_idx[e];
    const int off = __ldg(a.off + i), cnt = seg_cnt(a, i);
    const int b = off + c * SEG_HOT;
    const float4 acc = seg_sum_range(a, b, min(b + SEG_HOT, off + cnt), lane);
    st4(a.hot.chunk_sum + (long long)e * D + 4 * lane, acc);
  }
}

// kernel 3: one warp per hot segment adds its chunk sums in chunk order
static __global__ void __launch_bounds__(SEG_WARPS * 32) seg_hot_final_kernel(SegArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nhot = a.hot.ctr[0];
  for (int h = blockIdx.x * SEG_WARPS + warp; h < nhot; h += gridDim.x * SEG_WARPS) {
    const int i = a.hot.hot_seg[h], base = a.hot.hot_base[h];
    const int nch = (seg_cnt(a, i) + SEG_HOT - 1) / SEG_HOT;
    float4 acc = f4(0.f);
    for (int c = 0; c < nch; c += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (c + u < nch) ? ld4s(a.hot.chunk_sum + (long long)(base + c + u) * D + 4 * lane) : f4(0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u < nch) acc += v[u];
    }
    st4(a.out + (long long)i * D + 4 * lane, acc);
    const float sq = warp_sum(hsum4(acc * acc));
    int rank = 0;   // rank of this segment index among the hot ones: a deterministic slot for the sum of squares
    for (int j = lane; j < nhot; j += 32) rank += a.hot.hot_seg[j] < i;
    rank = (int)warp_sum((float)rank);
    if (lane == 0) a.hot.hot_sq[rank] = sq;
  }
}

// out[0] = sum(partials[0..n1)) + sum(hot_sq[0..n2)), fixed order, double accumulation
static __global__ void seg_sumsq_final_kernel(const float* __restrict__ p1, int n1, const float* __restrict__ p2, int n2, float* out) {
  __shared__ double s[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n1; i += 256) t += (double)p1[i];
  for (int i = threadIdx.x; i < n2; i += 256) t += (double)p2[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)s[0];
}

// Launches the three kernels (+ the clear of the hot work list).  grid1 = CTAs of the first kernel (sq_partial entries).
// n = total number of occurrences (bounds the hot lists).
inline cudaError_t seg_sum_launch(const SegArgs& a, int grid1, long long n, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(a.hot.ctr, 0, seg_clear_bytes(n), st);
  if (e != cudaSuccess) return e;
  // compact segment list over a table much larger than the batch: segments are short -> streaming variant
  if (a.nseg_dev != nullptr && a.nseg_bound >= 8 * n) seg_rows_kernel<8><<<grid1, SEG_WARPS * 32, 0, st>>>(a);
  else seg_rows_simple_kernel<<<grid1, SEG_WARPS * 32, 0, st>>>(a);
  const int g2 = (int)std::max<long long>(1, std::min<long long>((a.hot.max_chunks + SEG_WARPS - 1) / SEG_WARPS, 4LL * num_sms()));
  seg_hot_chunks_kernel<<<g2, SEG_WARPS * 32, 0, st>>>(a);
  const int g3 = (int)std::max<long long>(1, std::min<long long>((a.hot.max_hot + SEG_WARPS - 1) / SEG_WARPS, 2LL * num_sms()));
  seg_hot_final_kernel<<<g3, SEG_WARPS * 32, 0, st>>>(a);
  return cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace fbn

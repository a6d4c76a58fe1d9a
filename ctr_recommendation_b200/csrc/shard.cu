// Row-sharded item table: gradient exchange over NVLink peer memory + lazy row Adam (BASELINE config 5).
//
// The reference replicates item_emb on every GPU (nn.DataParallel, src/train_fibinet.py:69-70) and its backward is
// embedding_dense_backward + dense Adam (src/model_fibinet.py:159,167; src/train_fibinet.py:78).  Here global row g lives on
// rank g % N at local row g / N, and there is no all-to-all: the forward gather (embed.cu) loads rows straight from the
// owner's HBM, and the backward below lets every OWNER pull the partial gradient rows of its slice from all peers.
//
//   local stage  (per rank)  keys (owner*R + local row) -> stable radix sort -> run heads -> one warp per unique row sums
//                            its occurrences in source order -> exchange block {hdr, ukey[], ugrad[][128]} grouped by owner
//   merge stage  (per owner) copy the N key segments addressed to me (coalesced remote reads) -> rank every item among the
//                            N sorted lists by binary search (stable N-way merge: equal rows end up in rank order) ->
//                            run heads -> one warp per unique local row adds the <= N partial rows in rank order
//                            (512-byte remote loads) -> dense slice gradient + touched flags, or a compact (row, grad) list
//   update                   dense-exact Adam (optim.cu, fbn_adam_table on the slice) or lazy row Adam (adam_rows_kernel)
//
// Nothing here uses floating point atomics: every sum has a fixed order, so replicas and re-runs are bitwise reproducible.
#include <cub/cub.cuh>
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "segsum.cuh"
#include "tower.h"

namespace fbn {

constexpr int XHDR_BYTES = 256;     // int32 hdr[64]: [0] = U, [1 .. N+1] = owner_start[0..N]
constexpr int SH_WARPS = 8;

struct Xchg { int32_t* hdr; int32_t* ukey; float* ugrad; };

static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

__host__ __device__ inline Xchg xchg_view(void* base, long long cap) {
  Xchg x;
  char* p = static_cast<char*>(base);
  x.hdr = reinterpret_cast<int32_t*>(p);
  x.ukey = reinterpret_cast<int32_t*>(p + XHDR_BYTES);
  x.ugrad = reinterpret_cast<float*>(p + XHDR_BYTES + ((size_t(cap) * 4 + 255) & ~size_t(255)));
  return x;
}

struct ShardWs {
  // local stage
  int32_t* keys_in; int32_t* keys_out; int32_t* vals_in; int32_t* vals_out;
  int32_t* flags;   // (cap + 1)
  int32_t* uidx;    // (cap + 1) exclusive scan of flags
  int32_t* ustart;  // (cap + 1) first sorted occurrence of every unique row
  // merge stage
  int32_t* mhdr;    // [0] = T, [1] = Um, [2] = overflow flag, [8 + r] = s_r, [32 + r] = moff_r (r = 0..N)
  int32_t* mk;      // (merge_cap) local rows, concatenated by rank
  int32_t* mkey;    // (merge_cap + 1) merged order
  int32_t* mval;    // (merge_cap) peer * cap + index into the peer's ukey/ugrad
  int32_t* mflags;  // (merge_cap + 1)
  int32_t* midx;    // (merge_cap + 1)
  int32_t* mrow;    // (merge_cap) compact list: local row
  float* mgrad;     // (min(merge_cap, shard_rows), 128) compact list: summed gradient
  float* sq_partial;
  int sq_blocks;
  float* lsq_partial;  // per-CTA scratch of the local segment sums (their sum of squares is not used)
  int lsum_blocks;
  void* seg_scratch;   // hot-row work lists of segsum.cuh
  void* cub_tmp; size_t cub_bytes;
  size_t total;
};

static int key_bits_for(long long maxkey) {
  int b = 1;
  while ((1LL << b) <= maxkey) ++b;
  return b;
}

static int sum_blocks() { return 4 * num_sms(); }

static void carve_shard_ws(ShardWs& w, void* base, long long cap, long long merge_cap, int n, long long shard_rows) {
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? p + off : nullptr;
    off += al256(bytes);
    return r;
  };
  w.keys_in = (int32_t*)take(cap * 4); w.keys_out = (int32_t*)take(cap * 4);
  w.vals_in = (int32_t*)take(cap * 4); w.vals_out = (int32_t*)take(cap * 4);
  w.flags = (int32_t*)take((cap + 1) * 4); w.uidx = (int32_t*)take((cap + 1) * 4); w.ustart = (int32_t*)take((cap + 1) * 4);
  w.mhdr = (int32_t*)take(64 * 4);
  w.mk = (int32_t*)take(merge_cap * 4);
  w.mkey = (int32_t*)take((merge_cap + 1) * 4);
  w.mval = (int32_t*)take(merge_cap * 4);
  w.mflags = (int32_t*)take((merge_cap + 1) * 4);
  w.midx = (int32_t*)take((merge_cap + 1) * 4);
  w.mrow = (int32_t*)take(merge_cap * 4);
  w.mgrad = (float*)take((size_t)std::min(merge_cap, shard_rows) * D * sizeof(float));
  w.sq_blocks = sum_blocks();
  w.sq_partial = (float*)take((size_t)w.sq_blocks * sizeof(float));
  w.lsum_blocks = 16 * num_sms();
  w.lsq_partial = (float*)take((size_t)w.lsum_blocks * sizeof(float));
  w.seg_scratch = take(seg_scratch_bytes(cap));
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (int)std::max<long long>(cap, 1), 0, key_bits_for((long long)n * shard_rows));
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(cap + 1));
  cub::DeviceScan::ExclusiveSum(nullptr, c, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(merge_cap + 1));
  w.cub_bytes = std::max(a, std::max(b, c)) + 256;
  w.cub_tmp = take(w.cub_bytes);
  w.total = off;
}

// ---------------------------------------------------------------------------------------------------- local stage

__global__ void shard_keys_kernel(const void* __restrict__ item_id, int idx_dtype, const void* __restrict__ seq, int seq_dtype,
                                  long long B, int L, long long V, unsigned n, long long R, int32_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
  const long long cnt = B * (1 + (seq ? L : 0));
  const int sentinel = (int)((long long)n * R);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) {
    const long long raw = i < B ? load_index(item_id, idx_dtype, i) : load_index(seq, seq_dtype, i - B);
    const unsigned g = (unsigned)min(max(raw, 0LL), V - 1);   // same clamp as the forward gather
    keys[i] = g == 0 ? sentinel : (int)((long long)(g % n) * R + g / n);   // padding row 0 never receives a gradient
    vals[i] = (int)i;
  }
}

// flags[i] = 1 at the first element of every run of equal keys below `limit` (i < *n_dev or n_host); flags[i >= n] = 0
__global__ void run_heads_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ n_dev, long long n_host, long long total,
                                 int limit, int32_t* __restrict__ flags) {
  const long long n = n_dev ? (long long)*n_dev : n_host;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int f = 0;
    if (i < n) {
      const int k = keys[i];
      f = (k < limit) && (i == 0 || keys[i - 1] != k);
    }
    flags[i] = f;
  }
}

__global__ void shard_compact_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ flags, const int32_t* __restrict__ uidx,
                                     long long n, int32_t* __restrict__ ukey, int32_t* __restrict__ ustart) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (flags[i]) {
      const int u = uidx[i];
      ukey[u] = keys[i];
      ustart[u] = (int)i;
    }
}

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound_i32(const int32_t* a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// hdr[0] = U, hdr[1 + o] = first unique index owned by rank o (o = 0..N), ustart[U] = number of non-padding occurrences
__global__ void shard_hdr_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ uidx, int n, int nshards, long long R,
                                 int32_t* __restrict__ hdr, int32_t* __restrict__ ustart) {
  const int o = threadIdx.x;
  if (o <= nshards) {
    const int pos = lower_bound_i32(keys, n, (int)((long long)o * R));
    hdr[1 + o] = uidx[pos];
    if (o == nshards) {
      hdr[0] = uidx[n];
      ustart[uidx[n]] = pos;
    }
  }
}

// ---------------------------------------------------------------------------------------------------- merge stage

struct PeerPtrs { void* x[FBN_MAX_SHARDS]; };

// T = total number of partial rows addressed to me; s_r = start of peer r's segment, moff_r = its offset in the merged lists
__global__ void merge_plan_kernel(PeerPtrs peers, long long cap, int n, int me, long long merge_cap, int32_t* __restrict__ mhdr) {
  if (threadIdx.x == 0) {
    long long t = 0;
    int overflow = 0;
    for (int r = 0; r < n; ++r) {
      const int32_t* hdr = xchg_view(peers.x[r], cap).hdr;   // remote read (3 ints per peer)
      const int s = hdr[1 + me], e = hdr[2 + me];
      mhdr[8 + r] = s;
      mhdr[32 + r] = (int)t;
      long long len = e - s;
      if (t + len > merge_cap) { len = merge_cap - t; overflow = 1; }
      t += len;
    }
    mhdr[32 + n] = (int)t;
    mhdr[0] = (int)t;
    mhdr[2] = overflow;
  }
}

__device__ __forceinline__ int find_peer(const int32_t* moff, int n, int t) {
  int r = 0;
  while (r + 1 < n && t >= moff[r + 1]) ++r;
  return r;
}

// mk[t] = local row of the t-th partial row (peer segments concatenated in rank order); coalesced remote reads
__global__ void merge_copy_keys_kernel(PeerPtrs peers, long long cap, int n, int me, long long R, const int32_t* __restrict__ mhdr,
                                       int32_t* __restrict__ mk) {
  const int T = mhdr[0];
  const int32_t* moff = mhdr + 32;
  const int base = (int)((long long)me * R);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int r = find_peer(moff, n, t);
    const int j = mhdr[8 + r] + (t - moff[r]);
    mk[t] = xchg_view(peers.x[r], cap).ukey[j] - base;
  }
}

// stable N-way merge by ranking: position of item (r, i) = i + sum_{q<r} upper_bound_q(row) + sum_{q>r} lower_bound_q(row)
__global__ void merge_rank_kernel(long long cap, int n, const int32_t* __restrict__ mhdr, const int32_t* __restrict__ mk,
                                  int32_t* __restrict__ mkey, int32_t* __restrict__ mval) {
  const int T = mhdr[0];
  const int32_t* moff = mhdr + 32;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int r = find_peer(moff, n, t);
    const int row = mk[t];
    int pos = t - moff[r];
    for (int q = 0; q < n; ++q) {
      if (q == r) continue;
      const int len = moff[q + 1] - moff[q];
      pos += q < r ? upper_bound_i32(mk + moff[q], len, row) : lower_bound_i32(mk + moff[q], len, row);
    }
    mkey[pos] = row;
    mval[pos] = (int)(r * cap + mhdr[8 + r] + (t - moff[r]));
  }
}

// One warp per block of 32 merged items.  A warp owns the runs (equal local rows, <= N items, in rank order) whose head lies in
// its block and walks their items as one stream, MS_ILP at a time: MS_ILP independent 512-byte (remote) row loads in flight, then
// the adds in stream order with a store at every run end.  (One row per warp in flight was latency-bound: 437 GB/s over NVLink.)
constexpr int MS_ILP = 8;

__global__ void __launch_bounds__(SH_WARPS * 32) merge_sum_kernel(PeerPtrs peers, long long cap, int32_t* mhdr, long long merge_cap,
                                                                  const int32_t* __restrict__ mkey, const int32_t* __restrict__ mval,
                                                                  const int32_t* __restrict__ mflags, const int32_t* __restrict__ midx,
                                                                  float* __restrict__ dense_grad, int32_t* __restrict__ touched,
                                                                  int32_t* __restrict__ mrow, float* __restrict__ mgrad,
                                                                  float* __restrict__ sq_partial) {
  __shared__ float s_sq[SH_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = mhdr[0];
  const int icap = (int)cap;
  float sq = 0.f;
  for (long long t0 = ((long long)blockIdx.x * SH_WARPS + warp) * 32; t0 < T; t0 += (long long)gridDim.x * SH_WARPS * 32) {
    // this block and a 32-item look-ahead (a run never exceeds FBN_MAX_SHARDS = 16 items)
    const long long ta = t0 + lane, tb = t0 + 32 + lane;
    const int f0 = ta < T ? mflags[ta] : 0, f1 = tb < T ? mflags[tb] : 0;
    const int k0 = ta < T ? mkey[ta] : 0;
    const int v0 = ta < T ? mval[ta] : 0, v1 = tb < T ? mval[tb] : 0;
    const int x0 = ta < T ? midx[ta] : 0;
    const unsigned h0 = __ballot_sync(0xffffffffu, f0 != 0), h1 = __ballot_sync(0xffffffffu, f1 != 0);
    if (h0 == 0) continue;                                   // (only possible past the end of the list)
    const int p_begin = __ffs(h0) - 1;
    const int p_end = h1 ? 32 + __ffs(h1) - 1 : (int)min((long long)64, (long long)T - t0);
    auto flag_at = [&](int rel) { return rel < 32 ? (int)((h0 >> rel) & 1u) : (int)((h1 >> (rel - 32)) & 1u); };
    int head_rel = p_begin;
    float4 acc = f4(0.f);
    for (int p = p_begin; p < p_end; p += MS_ILP) {
      float4 v[MS_ILP];
#pragma unroll
      for (int u = 0; u < MS_ILP; ++u) {
        const int rel = p + u;
        const int val = rel < 32 ? __shfl_sync(0xffffffffu, v0, rel & 31) : __shfl_sync(0xffffffffu, v1, rel & 31);
        if (rel < p_end) {
          const int r = val / icap;
          const long long j = val - r * icap;
          v[u] = ld4s(xchg_view(peers.x[r], cap).ugrad + j * D + 4 * lane);
        } else {
          v[u] = f4(0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < MS_ILP; ++u) {
        const int rel = p + u;
        if (rel < p_end) {
          if (flag_at(rel)) { head_rel = rel; acc = f4(0.f); }
          acc += v[u];
          if (rel + 1 == p_end || flag_at(rel + 1)) {        // run complete: rel is its last item
            const int row = __shfl_sync(0xffffffffu, k0, head_rel & 31);   // heads of my runs are always inside my block
            if (dense_grad) {
              st4(dense_grad + (long long)row * D + 4 * lane, acc);
              if (lane == 0) touched[row] = 1;
            } else {
              const int um = __shfl_sync(0xffffffffu, x0, head_rel & 31);
              st4(mgrad + (long long)um * D + 4 * lane, acc);
              if (lane == 0) mrow[um] = row;
            }
            sq += warp_sum(hsum4(acc * acc));
          }
        }
      }
    }
  }
  if (lane == 0) s_sq[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tt = 0.f;
#pragma unroll
    for (int w = 0; w < SH_WARPS; ++w) tt += s_sq[w];
    sq_partial[blockIdx.x] = tt;
    if (blockIdx.x == 0) mhdr[1] = midx[merge_cap];   // Um
  }
}

__global__ void shard_sum_partials_kernel(const float* __restrict__ x, int n, float* out) {
  __shared__ double s[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += (double)x[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)s[0];
}

// ---------------------------------------------------------------------------------------------------- lazy row Adam

struct RowAdamHyper { float lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, omb1, omb2, decay_mul; };

__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, const RowAdamHyper& h) {
  if (h.decay_mul != 0.f) p *= h.decay_mul;          // same op order as adam1 (optim.cu)
  else g = fmaf(h.wd, p, g);
  m = m + h.omb1 * (g - m);
  v = v * h.beta2 + h.omb2 * (g * g);
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = p - h.step_size * (m / denom);
}

__global__ void __launch_bounds__(SH_WARPS * 32) adam_rows_kernel(const int32_t* __restrict__ mhdr, const int32_t* __restrict__ mrow,
                                                                  const float* __restrict__ mgrad, float* __restrict__ p,
                                                                  float* __restrict__ m, float* __restrict__ v,
                                                                  const float* __restrict__ clip, RowAdamHyper hv,
                                                                  const float* __restrict__ hdev) {
  RowAdamHyper h = hv;
  if (hdev) { h.lr = hdev[0]; h.beta1 = hdev[1]; h.beta2 = hdev[2]; h.eps = hdev[3]; h.wd = hdev[4]; h.step_size = hdev[5]; h.bc2_sqrt = hdev[6]; h.omb1 = hdev[8]; h.omb2 = hdev[9]; h.decay_mul = hdev[10]; }
  const float coef = clip ? clip[1] : 1.0f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Um = mhdr[1];
  for (long long u = (long long)blockIdx.x * SH_WARPS + warp; u < Um; u += (long long)gridDim.x * SH_WARPS) {
    const long long e = (long long)mrow[u] * D + 4 * lane;
    const float4 g = ld4s(mgrad + u * D + 4 * lane) * coef;
    float4 pp = ld4s(p + e), mm = ld4s(m + e), vv = ld4s(v + e);
    adam_elem(pp.x, mm.x, vv.x, g.x, h); adam_elem(pp.y, mm.y, vv.y, g.y, h);
    adam_elem(pp.z, mm.z, vv.z, g.z, h); adam_elem(pp.w, mm.w, vv.w, g.w, h);
    st4(p + e, pp); st4(m + e, mm); st4(v + e, vv);
  }
}

}  // namespace fbn

using namespace fbn;

static int check_plan(const fbn_shard_plan_t* s, void* sws, size_t sws_bytes, ShardWs& w) {
  FBN_REQUIRE(s && sws, FBN_ERR_ARG, "fbn_shard: null plan / workspace");
  FBN_REQUIRE(s->n_shards >= 1 && s->n_shards <= FBN_MAX_SHARDS && s->rank >= 0 && s->rank < s->n_shards, FBN_ERR_ARG,
              "fbn_shard: n_shards must be in [1,%d] and rank in [0,n_shards)", FBN_MAX_SHARDS);
  FBN_REQUIRE(s->item_rows >= 2 && s->shard_rows == cdiv(s->item_rows, s->n_shards), FBN_ERR_SHAPE,
              "fbn_shard: shard_rows must be ceil(item_rows / n_shards)");
  FBN_REQUIRE(s->cap >= 1 && s->merge_cap >= 1 && s->merge_cap <= (long long)s->n_shards * s->cap, FBN_ERR_SHAPE, "fbn_shard: bad capacities");
  FBN_REQUIRE((long long)s->n_shards * s->shard_rows < (1LL << 31) - 1 && (long long)s->n_shards * s->cap < (1LL << 31), FBN_ERR_SHAPE,
              "fbn_shard: table or batch too large for 32-bit keys");
  for (int r = 0; r < s->n_shards; ++r) FBN_REQUIRE(s->xchg[r] && aligned16(s->xchg[r]), FBN_ERR_ALIGN, "fbn_shard: exchange block %d missing / unaligned", r);
  FBN_REQUIRE(aligned16(sws), FBN_ERR_ALIGN, "fbn_shard: workspace is not 16-byte aligned");
  carve_shard_ws(w, sws, s->cap, s->merge_cap, s->n_shards, s->shard_rows);
  FBN_REQUIRE(sws_bytes >= w.total, FBN_ERR_ARG, "fbn_shard: workspace too small: %zu < %zu", sws_bytes, w.total);
  return FBN_OK;
}

extern "C" size_t fbn_shard_xchg_bytes(int64_t cap) { return XHDR_BYTES + al256((size_t)cap * 4) + (size_t)cap * D * sizeof(float); }

extern "C" size_t fbn_shard_ws_bytes(int64_t cap, int64_t merge_cap, int n_shards, int64_t shard_rows) {
  ShardWs w;
  carve_shard_ws(w, nullptr, cap, merge_cap, n_shards, shard_rows);
  return w.total;
}

static int grid_for(long long n, int threads) { return (int)std::max<long long>(1, std::min<long long>(cdiv(n, threads), 8LL * num_sms())); }

extern "C" int fbn_shard_index(const fbn_shard_plan_t* s, const fbn_batch_t* b, void* sws, size_t sws_bytes, fbn_stream_t stream) {
  ShardWs w;
  int rc = check_plan(s, sws, sws_bytes, w);
  if (rc) return rc;
  FBN_REQUIRE(b && b->item_id && b->batch >= 1, FBN_ERR_ARG, "fbn_shard_index: bad batch");
  cudaStream_t st = (cudaStream_t)stream;
  const void* seq = (b->seq_len > 0 && b->item_seq) ? b->item_seq : nullptr;
  const long long n = b->batch * (1 + (seq ? b->seq_len : 0));
  FBN_REQUIRE(n <= s->cap, FBN_ERR_SHAPE, "fbn_shard_index: %lld occurrences exceed the exchange capacity %lld", n, (long long)s->cap);
  const long long R = s->shard_rows;
  const int sentinel = (int)((long long)s->n_shards * R);
  Xchg x = xchg_view(s->xchg[s->rank], s->cap);
  shard_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(b->item_id, b->idx_dtype, seq, b->seq_dtype, b->batch, (int)b->seq_len, s->item_rows,
                                                      (unsigned)s->n_shards, R, w.keys_in, w.vals_in);
  FBN_CHECK_LAUNCH();
  size_t bytes = w.cub_bytes;
  FBN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, (const int32_t*)w.keys_in, w.keys_out, (const int32_t*)w.vals_in,
                                                 w.vals_out, (int)n, 0, key_bits_for(sentinel), st));
  g_launches += 4;
  run_heads_kernel<<<grid_for(n + 1, 256), 256, 0, st>>>(w.keys_out, nullptr, n, n + 1, sentinel, w.flags);
  FBN_CHECK_LAUNCH();
  bytes = w.cub_bytes;
  FBN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, (const int32_t*)w.flags, w.uidx, (int)(n + 1), st));
  g_launches += 2;
  shard_compact_kernel<<<grid_for(n, 256), 256, 0, st>>>(w.keys_out, w.flags, w.uidx, n, x.ukey, w.ustart);
  FBN_CHECK_LAUNCH();
  shard_hdr_kernel<<<1, 32, 0, st>>>(w.keys_out, w.uidx, (int)n, s->n_shards, R, x.hdr, w.ustart);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_shard_local_sum(const fbn_shard_plan_t* s, const fbn_batch_t* b, const float* dXitem, const float* dXhist, void* sws,
                                   size_t sws_bytes, fbn_stream_t stream) {
  ShardWs w;
  int rc = check_plan(s, sws, sws_bytes, w);
  if (rc) return rc;
  FBN_REQUIRE(b && dXitem && dXhist && aligned16(dXitem) && aligned16(dXhist), FBN_ERR_ARG, "fbn_shard_local_sum: bad arguments");
  const void* seq = (b->seq_len > 0 && b->item_seq) ? b->item_seq : nullptr;
  const long long n = b->batch * (1 + (seq ? b->seq_len : 0));
  Xchg x = xchg_view(s->xchg[s->rank], s->cap);
  // one warp per unique row sums its occurrences in source order; hot rows are chunked (segsum.cuh) -- the same value,
  // bit for bit, as the replicated table's emb_rows for that row
  SegArgs g{};
  g.off = w.ustart; g.cnt = nullptr; g.nseg_dev = x.hdr; g.nseg = 0; g.src = w.vals_out;
  g.dXitem = dXitem; g.dXhist = dXhist; g.B = b->batch; g.L = b->seq_len > 0 ? (int)b->seq_len : 1;
  g.out = x.ugrad; g.zero_fill = 0; g.sq_partial = w.lsq_partial; g.nseg_bound = s->item_rows;
  g.hot = seg_carve(w.seg_scratch, s->cap);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n, SEG_WARPS), w.lsum_blocks));
  FBN_CHECK_CUDA(seg_sum_launch(g, blocks, s->cap, (cudaStream_t)stream));
  g_launches += 3;
  return FBN_OK;
}

extern "C" int fbn_shard_merge(const fbn_shard_plan_t* s, void* sws, size_t sws_bytes, float* dense_grad, int32_t* touched,
                               float* sumsq_out, fbn_stream_t stream) {
  ShardWs w;
  int rc = check_plan(s, sws, sws_bytes, w);
  if (rc) return rc;
  FBN_REQUIRE((dense_grad == nullptr) == (touched == nullptr), FBN_ERR_ARG, "fbn_shard_merge: dense_grad and touched go together");
  FBN_REQUIRE(sumsq_out != nullptr && aligned16(dense_grad), FBN_ERR_ARG, "fbn_shard_merge: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  PeerPtrs peers{};
  for (int r = 0; r < s->n_shards; ++r) peers.x[r] = s->xchg[r];
  const long long mc = s->merge_cap;
  if (touched) FBN_CHECK_CUDA(cudaMemsetAsync(touched, 0, sizeof(int32_t) * s->shard_rows, st));
  merge_plan_kernel<<<1, 32, 0, st>>>(peers, s->cap, s->n_shards, s->rank, mc, w.mhdr);
  FBN_CHECK_LAUNCH();
  merge_copy_keys_kernel<<<grid_for(mc, 256), 256, 0, st>>>(peers, s->cap, s->n_shards, s->rank, s->shard_rows, w.mhdr, w.mk);
  FBN_CHECK_LAUNCH();
  merge_rank_kernel<<<grid_for(mc, 256), 256, 0, st>>>(s->cap, s->n_shards, w.mhdr, w.mk, w.mkey, w.mval);
  FBN_CHECK_LAUNCH();
  run_heads_kernel<<<grid_for(mc + 1, 256), 256, 0, st>>>(w.mkey, w.mhdr, 0, mc + 1, 0x7fffffff, w.mflags);
  FBN_CHECK_LAUNCH();
  size_t bytes = w.cub_bytes;
  FBN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, (const int32_t*)w.mflags, w.midx, (int)(mc + 1), st));
  g_launches += 2;
  merge_sum_kernel<<<w.sq_blocks, SH_WARPS * 32, 0, st>>>(peers, s->cap, w.mhdr, mc, w.mkey, w.mval, w.mflags, w.midx, dense_grad,
                                                         touched, w.mrow, w.mgrad, w.sq_partial);
  FBN_CHECK_LAUNCH();
  shard_sum_partials_kernel<<<1, 256, 0, st>>>(w.sq_partial, w.sq_blocks, sumsq_out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_shard_adam_rows(const fbn_shard_plan_t* s, void* sws, size_t sws_bytes, float* p, float* m, float* v, const float* clip,
                                   const fbn_adam_t* h, const float* hyper_dev, fbn_stream_t stream) {
  ShardWs w;
  int rc = check_plan(s, sws, sws_bytes, w);
  if (rc) return rc;
  FBN_REQUIRE(p && m && v && (h || hyper_dev), FBN_ERR_ARG, "fbn_shard_adam_rows: null pointer");
  FBN_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v), FBN_ERR_ALIGN, "fbn_shard_adam_rows: unaligned pointer");
  RowAdamHyper hv{};
  if (h) {
    FBN_REQUIRE(h->step >= 1, FBN_ERR_ARG, "fbn_shard_adam_rows: step must be >= 1");
    hv.lr = h->lr; hv.beta1 = h->beta1; hv.beta2 = h->beta2; hv.eps = h->eps; hv.wd = h->weight_decay;
    hv.step_size = (float)((double)h->lr / (1.0 - pow((double)h->beta1, (double)h->step)));
    hv.bc2_sqrt = (float)sqrt(1.0 - pow((double)h->beta2, (double)h->step));
    hv.omb1 = h->one_minus_beta1 > 0.f ? h->one_minus_beta1 : one_minus(h->beta1);
    hv.omb2 = h->one_minus_beta2 > 0.f ? h->one_minus_beta2 : one_minus(h->beta2);
    hv.decay_mul = h->decoupled ? (float)(1.0 - (double)h->lr * (double)h->weight_decay) : 0.f;
  }
  const long long upper = std::min<long long>(s->merge_cap, s->shard_rows);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(upper, SH_WARPS), 8LL * num_sms()));
  adam_rows_kernel<<<blocks, SH_WARPS * 32, 0, (cudaStream_t)stream>>>(w.mhdr, w.mrow, w.mgrad, p, m, v, clip, hv, hyper_dev);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_shard_stats(const fbn_shard_plan_t* s, void* sws, size_t sws_bytes, int32_t* host_out, fbn_stream_t stream) {
  ShardWs w;
  int rc = check_plan(s, sws, sws_bytes, w);
  if (rc) return rc;
  FBN_REQUIRE(host_out, FBN_ERR_ARG, "fbn_shard_stats: null output");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t hdr[64], mh[64];
  FBN_CHECK_CUDA(cudaMemcpyAsync(hdr, xchg_view(s->xchg[s->rank], s->cap).hdr, sizeof(hdr), cudaMemcpyDeviceToHost, st));
  FBN_CHECK_CUDA(cudaMemcpyAsync(mh, w.mhdr, sizeof(mh), cudaMemcpyDeviceToHost, st));
  FBN_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 24; ++i) host_out[i] = 0;
  host_out[0] = hdr[0];
  for (int o = 0; o <= s->n_shards; ++o) host_out[1 + o] = hdr[1 + o];
  host_out[20] = mh[0]; host_out[21] = mh[1]; host_out[22] = mh[2];
  return FBN_OK;
}

// ---------------------------------------------------------------------------------------------------- CUDA IPC plumbing

extern "C" int fbn_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  FBN_REQUIRE(dev_ptr && handle_out && offset_out, FBN_ERR_ARG, "fbn_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  FBN_CHECK_CUDA(cudaFree(nullptr));   // make sure the primary context is current
  // resolved through the runtime so that the library carries no link-time dependency on libcuda.so.1
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range_fn = nullptr;
  if (!range_fn) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    FBN_CHECK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q));
    FBN_REQUIRE(q == cudaDriverEntryPointSuccess && fp, FBN_ERR_CUDA, "cuMemGetAddressRange is not available from the driver");
    range_fn = reinterpret_cast<RangeFn>(fp);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult cr = range_fn(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr);
  FBN_REQUIRE(cr == CUDA_SUCCESS, FBN_ERR_CUDA, "cuMemGetAddressRange failed (%d)", (int)cr);
  FBN_CHECK_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out), reinterpret_cast<void*>(base)));
  *offset_out = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  return FBN_OK;
}

extern "C" int fbn_ipc_open(const void* handle, void** base_out) {
  FBN_REQUIRE(handle && base_out, FBN_ERR_ARG, "fbn_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  FBN_CHECK_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return FBN_OK;
}

extern "C" int fbn_ipc_close(void* base) {
  FBN_REQUIRE(base, FBN_ERR_ARG, "fbn_ipc_close: null pointer");
  FBN_CHECK_CUDA(cudaIpcCloseMemHandle(base));
  return FBN_OK;
}

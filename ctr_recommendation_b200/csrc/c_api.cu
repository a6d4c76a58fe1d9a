// extern "C" surface of libfibinet_b200.so: argument validation + the launch sequence of one
// forward / backward pass.  All orchestration is native so that a training step costs one FFI call
// per phase (and can be captured into a CUDA graph by the host).
#include <algorithm>
#include <stdarg.h>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "gemm.h"
#include "tower.h"

namespace fbn {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else {
      cudaGetLastError();
      sms = 148;  // B200
    }
  }
  return sms;
}

static int pick_splits(long long tiles, long long K) {
  const long long ktiles = std::max<long long>(1, cdiv(K, 16));
  long long s = cdiv(2LL * 148, std::max<long long>(tiles, 1));
  s = std::min<long long>(s, 32);
  s = std::min<long long>(s, ktiles);
  return (int)std::max<long long>(s, 1);
}

// K blocks of the MLP input that can be non-zero: fields 1..5 and pairs (i>=1, j)
static unsigned long long active_mask() {
  unsigned long long m = 0;
  for (int f = 1; f < NF; ++f) m |= 1ull << f;
  for (int blk = NF + (NF - 1); blk < NF + FBN_PAIRS; ++blk) m |= 1ull << blk;
  return m;
}

static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

void carve_workspace(Workspace& w, void* base, int64_t B, int64_t L, int64_t rows) {
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? p + off : nullptr;
    off += align_up(bytes);
    return r;
  };
  const size_t f = sizeof(float);
  const int64_t Bp = std::max<int64_t>(B, 1);
  const int64_t Lp = std::max<int64_t>(L, 1);
  w.B = B; w.L = L;
  w.ids = (int32_t*)take(Bp * 4 * 4);
  w.seq = (int32_t*)take(Bp * Lp * 4);
  w.idflag = (int32_t*)take(4 * 4);
  w.X5 = (float*)take(Bp * NA * D * f);
  w.sgate = (float*)take(Bp * 8 * f);
  w.xhat = (float*)take(Bp * D * f);
  w.xmm = (float*)take(Bp * D * f);
  w.Ymm = (float*)take(Bp * D * f);
  w.rstd = (float*)take(Bp * f);
  w.cnt = (float*)take(Bp * f);
  w.C = (float*)take(Bp * K1 * f);
  w.T = (float*)take(Bp * 10 * D * f);
  w.Hd1 = (float*)take(Bp * H1 * f);
  w.A1 = (float*)take(Bp * H1 * f);
  w.Hd2 = (float*)take(Bp * H2 * f);
  w.A2 = (float*)take(Bp * H2 * f);
  w.logit = (float*)take(Bp * f);
  w.prob = (float*)take(Bp * f);
  w.bn = (float*)take((2 * H1 + 2 * H2) * f);
  w.dlogit = (float*)take(Bp * f);
  w.dH2 = (float*)take(Bp * H2 * f);
  w.dH1 = (float*)take(Bp * H1 * f);
  w.dC = (float*)take(Bp * K1 * f);
  w.dT = (float*)take(Bp * 10 * D * f);
  w.dV = (float*)take(Bp * NA * D * f);
  w.dXitem = (float*)take(Bp * D * f);
  w.dXhist = (float*)take(Bp * D * f);
  w.dln = (float*)take(Bp * D * f);
  w.dy = (float*)take(Bp * D * f);
  w.sestat = (float*)take(Bp * 24 * f);
  size_t pf = 0;
  pf = std::max<size_t>(pf, (size_t)12 * H1 * K1);  // up to 12-way split-K of the MLP-1 weight gradient
  pf = std::max<size_t>(pf, (size_t)pick_splits(2 * 4, Bp) * H2 * H1);
  pf = std::max<size_t>(pf, (size_t)10 * 32 * D * D);
  pf = std::max<size_t>(pf, (size_t)2 * 148 * 4 * MAX_CATE * D);
  pf = std::max<size_t>(pf, (size_t)16 * 148 * 3 * H1);
  pf = std::max<size_t>(pf, (size_t)cdiv(rows, 8) + 1024);
  pf = std::max<size_t>(pf, (size_t)1024 * 96);
  w.partial_floats = pf;
  w.partial = (float*)take(pf * f);
  w.partial_side = (float*)take(pf * f);
  w.partial_cate = (float*)take((size_t)2 * 148 * 4 * MAX_CATE * D * f);
  w.partial_embsq = (float*)take(((size_t)cdiv(rows, 8) + 1024) * f);
  const int64_t nocc = Bp * (1 + Lp);
  w.keys_in = (int32_t*)take(nocc * 4);
  w.keys_out = (int32_t*)take(nocc * 4);
  w.vals_in = (int32_t*)take(nocc * 4);
  w.vals_out = (int32_t*)take(nocc * 4);
  w.row_off = (int32_t*)take((rows + 1) * 4);
  w.row_cnt = (int32_t*)take((rows + 1) * 4);
  w.cub_bytes = emb_sort_temp_bytes(nocc, rows);
  w.cub_tmp = take(w.cub_bytes);
  // operand scratch of the tcgen05 path: the largest call is the MLP-1 weight gradient (512 + 2688) x B
  size_t gs = 4096;   // every operand of the model path is pre-packed into the pk_* regions below
  w.gemm_scratch_bytes = gs;
  w.gemm_scratch = take(gs);
  // each region is sized for the larger of the formats it may hold (tf32 hi|lo: 8 bytes / element; f16x3: 4 + its scale record)
  auto pkb = [](long long r, long long c) { return std::max(packed_bytes(r, c, FBN_PREC_TF32X3), packed_bytes(r, c, FBN_PREC_F16X3)); };
  w.pk_C = take(pkb(Bp, K1));
  w.pk_A1 = take(pkb(Bp, H1));
  w.pk_dH2 = take(pkb(Bp, H2));
  w.pk_dH1 = take(pkb(Bp, H1));
  w.pk_dT = take(pkb(Bp, 10 * D));
  w.pk_dy = take(pkb(Bp, D));
  w.pk_xmm = take(pkb(Bp, D));
  w.pk_w1 = take(pkb(H1, K1));
  w.pk_w2 = take(pkb(H2, H1));
  w.pk_bil = take(pkb(FBN_PAIRS * D, D));
  w.pk_mmw = take(pkb(D, D));
  // f16x3: the MLP input is needed twice -- its field blocks as tf32 hi|lo (pk_C, written by the gather kernel for the bilinear
  // transforms) and the whole live row as fp16 hi|lo under one scale (the MLP-1 forward / weight-gradient GEMMs)
  w.pk16_C = take(packed_bytes(Bp, K1, FBN_PREC_F16X3));
  w.total_bytes = off;
}

int gemm(const GemmArgs& g, int precision, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  if (precision == FBN_PREC_FP32) return gemm_simt(g, st);
  return gemm_tc(g, precision, scratch, scratch_bytes, st);
}

// Registry of the tensors packed so far in this step: a GEMM operand given as an fp32 pointer anywhere inside a
// registered (contiguous) tensor is served from its packed copy at the same element offset (pitch == ld for all of
// them), so each activation / weight is converted exactly once per step.
struct PkReg {
  int prec = FBN_PREC_FP32;
  cudaStream_t st = nullptr;
  struct E { const float* src; size_t n; Packed pk; int fmt; } e[24];
  int ne = 0;
  bool on() const { return prec != FBN_PREC_FP32; }
  // FBN_PREC_F16X3 is a two-format mode: the operands of the long-K MLP GEMMs are fp16 hi|lo under a per-tensor scale (found by
  // an amax pass over the finished tensor, so they are packed AFTER their producer), everything the short-K GEMMs read keeps
  // the tf32 hi|lo format written by the producer kernels themselves
  bool mlp16() const { return prec == FBN_PREC_F16X3; }
  int lowfmt() const { return mlp16() ? FBN_PREC_TF32X3 : prec; }
  int mlpfmt() const { return prec; }
  static int esz(int fmt) { return fmt == FBN_PREC_TF32X3 ? 4 : 2; }
  void describe(const float* src, long long rows, long long cols, void* region, int fmt = -1) {
    if (!on()) return;
    if (fmt < 0) fmt = lowfmt();
    for (int i = 0; i < ne; ++i) if (e[i].src == src && e[i].fmt == fmt) return;
    e[ne].src = src; e[ne].n = (size_t)rows * cols; e[ne].pk = packed_describe(region, rows, cols, fmt); e[ne].fmt = fmt;
    ++ne;
  }
  int pack(const float* src, long long rows, long long cols, void* region, unsigned long long colmask = ~0ull, int fmt = -1) {
    if (!on()) return FBN_OK;
    if (fmt < 0) fmt = lowfmt();
    describe(src, rows, cols, region, fmt);
    Packed tmp;
    return pack_operand(src, cols, rows, cols, fmt, region, colmask, &tmp, st);
  }
  // an operand of the MLP GEMMs only (weights, or an activation after its producer when the producer could not pack it)
  int pack_mlp(const float* src, long long rows, long long cols, void* region, unsigned long long colmask = ~0ull) {
    return pack(src, rows, cols, region, colmask, mlpfmt());
  }
  void describe_mlp(const float* src, long long rows, long long cols, void* region) { describe(src, rows, cols, region, mlpfmt()); }
  // register `src` as packed-by-its-producer and return the destination descriptor for that kernel
  PackDst dst(const float* src, long long rows, long long cols, void* region) {
    PackDst d;
    if (!on()) return d;
    describe(src, rows, cols, region, lowfmt());
    const Packed pk = packed_describe(region, rows, cols, lowfmt());
    d.base = pk.data; d.pitch = pk.pitch; d.lo_off = pk.lo_off; d.mode = lowfmt();
    return d;
  }
  // the same for a tensor only the MLP GEMMs read: in f16x3 mode the producer writes fp32 only and the caller follows up with
  // after_mlp() once the tensor is complete
  PackDst dst_mlp(const float* src, long long rows, long long cols, void* region) {
    if (!mlp16()) return dst(src, rows, cols, region);
    PackDst d;          // mode 0: fp32 only, plus the per-block maxima the split pass needs
    d.tail = packed_describe(region, rows, cols, FBN_PREC_F16X3).scale;
    return d;
  }
  int after_mlp(const float* src, long long rows, long long cols, void* region) {
    if (!mlp16()) return FBN_OK;
    describe(src, rows, cols, region, FBN_PREC_F16X3);
    Packed tmp;
    return pack_operand(src, cols, rows, cols, FBN_PREC_F16X3, region, ~0ull, &tmp, st, /*producer_amax=*/true);
  }
  Packed find(const float* p, int fmt) const {
    for (int i = 0; i < ne; ++i)
      if (e[i].fmt == fmt && p >= e[i].src && p < e[i].src + e[i].n) return e[i].pk.view_cols((long long)(p - e[i].src), esz(fmt));
    return Packed();
  }
  int run(GemmArgs g, Workspace& w) const { return run_on(g, w, st); }
  int run_on(GemmArgs g, Workspace& w, cudaStream_t s) const {
    int fmt = prec;
    if (on()) {
      fmt = lowfmt();
      if (mlp16()) {      // both operands available as fp16 hi|lo -> the f16x3 kernel, otherwise the tf32x3 one
        const Packed a16 = find(g.A, FBN_PREC_F16X3), b16 = find(g.B, FBN_PREC_F16X3);
        if (a16.data && b16.data) { g.pkA = a16; g.pkB = b16; return gemm(g, FBN_PREC_F16X3, w.gemm_scratch, w.gemm_scratch_bytes, s); }
      }
      g.pkA = find(g.A, fmt); g.pkB = find(g.B, fmt);
    }
    return gemm(g, fmt, w.gemm_scratch, w.gemm_scratch_bytes, s);
  }
};
static thread_local PkReg tl_reg;   // rebuilt at the start of every fbn_forward / fbn_backward call

// Optional stage timing (fbn_set_option("stage_events", 1), eager launches only): a CUDA event is recorded on the calling
// stream at every stage boundary of fbn_forward / fbn_backward; fbn_stage_report turns them into per-stage milliseconds.
struct StageLog { bool on = false; std::vector<std::pair<std::string, cudaEvent_t>> ev; };
static StageLog g_stage;
static int stage_mark(const char* name, cudaStream_t st) {
  if (!g_stage.on) return FBN_OK;
  cudaEvent_t e;
  FBN_CHECK_CUDA(cudaEventCreate(&e));
  FBN_CHECK_CUDA(cudaEventRecord(e, st));
  g_stage.ev.emplace_back(name, e);
  return FBN_OK;
}
void set_stage_events(int on) { g_stage.on = on != 0; }

// launchers defined in embed.cu
struct EmbedFwdArgs;
struct EmbedBwdArgs;

}  // namespace fbn

#include "embed_args.h"

using namespace fbn;

#define STAGE(name) RC(stage_mark(name, st))
#define RC(x)            \
  do {                   \
    int _rc = (x);       \
    if (_rc) return _rc; \
  } while (0)

// rows the workspace is sized for: the dense per-row index of the replicated table does not exist in row-sharded mode
static inline int64_t ws_rows(const fbn_params_t* p) { return p->n_shards > 0 ? 1 : p->item_rows; }

static int check_common(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes) {
  FBN_REQUIRE(p && b && ws, FBN_ERR_ARG, "null params / batch / workspace");
  FBN_REQUIRE(p->n_shards >= 0 && p->n_shards <= FBN_MAX_SHARDS, FBN_ERR_ARG, "n_shards must be in [0,%d]", FBN_MAX_SHARDS);
  if (p->n_shards > 0) {
    FBN_REQUIRE(p->shard_rank >= 0 && p->shard_rank < p->n_shards && p->shard_rows * p->n_shards >= p->item_rows &&
                    p->item_rows < (1LL << 31), FBN_ERR_SHAPE, "inconsistent shard description");
    for (int r = 0; r < p->n_shards; ++r)
      FBN_REQUIRE(p->shard[r] && aligned16(p->shard[r]), FBN_ERR_ALIGN, "shard pointer %d missing / unaligned", r);
  }
  FBN_REQUIRE(b->batch >= 1, FBN_ERR_SHAPE, "batch must be >= 1");
  FBN_REQUIRE(b->seq_len >= 0 && b->seq_len <= MAX_L, FBN_ERR_SHAPE, "seq_len must be in [0,%d]", MAX_L);
  FBN_REQUIRE(p->cate_rows >= 1 && p->cate_rows <= MAX_CATE, FBN_ERR_SHAPE, "cate_rows must be in [1,%d]", MAX_CATE);
  FBN_REQUIRE(p->item_rows >= 2, FBN_ERR_SHAPE, "item_rows must be >= 2");
  FBN_REQUIRE(b->item_id && b->likes_level && b->views_level, FBN_ERR_ARG, "missing index columns");
  FBN_REQUIRE(b->item_mm || b->mm_table, FBN_ERR_ARG, "item_emb_d128 (item_mm) or mm_table is required");
  FBN_REQUIRE(b->idx_dtype >= FBN_IDX_I32 && b->idx_dtype <= FBN_IDX_F32, FBN_ERR_DTYPE, "bad idx_dtype");
  FBN_REQUIRE(b->seq_dtype == FBN_IDX_I32 || b->seq_dtype == FBN_IDX_I64, FBN_ERR_DTYPE, "item_seq must be int32 or int64");
  FBN_REQUIRE(p->bilinear_type >= FBN_BILINEAR_ALL && p->bilinear_type <= FBN_BILINEAR_INTERACTION, FBN_ERR_ARG, "bad bilinear_type");
  FBN_REQUIRE((p->precision >= FBN_PREC_FP32 && p->precision <= FBN_PREC_BF16) || p->precision == FBN_PREC_F16X3, FBN_ERR_ARG, "bad precision");
  const void* ptrs[] = {p->item_emb, p->cate_emb, p->mm_w, p->mm_b, p->ln_g, p->ln_b, p->bil_w, p->w1, p->b1, p->bn1_g, p->bn1_b,
                        p->bn1_mean, p->bn1_var, p->w2, p->b2, p->bn2_g, p->bn2_b, p->bn2_mean, p->bn2_var, p->w3, ws,
                        b->item_mm, b->mm_table};
  for (const void* q : ptrs) FBN_REQUIRE(aligned16(q), FBN_ERR_ALIGN, "a tensor pointer is not 16-byte aligned");
  FBN_REQUIRE(p->se_w1 && p->se_b1 && p->se_w2 && p->se_b2 && p->b3, FBN_ERR_ARG, "null parameter pointer");
  Workspace w;
  carve_workspace(w, nullptr, b->batch, b->seq_len, ws_rows(p));
  FBN_REQUIRE(ws_bytes >= w.total_bytes, FBN_ERR_ARG, "workspace too small: %zu < %zu", ws_bytes, w.total_bytes);
  return FBN_OK;
}

extern "C" size_t fbn_workspace_bytes(int64_t batch, int64_t seq_len, int64_t item_rows) {
  Workspace w;
  carve_workspace(w, nullptr, batch, seq_len, item_rows);
  return w.total_bytes;
}

extern "C" size_t fbn_workspace_offset(int64_t batch, int64_t seq_len, int64_t item_rows, const char* name) {
  Workspace w;
  char* base = reinterpret_cast<char*>(uintptr_t(4096));
  carve_workspace(w, base, batch, seq_len, item_rows);
  struct { const char* n; void* p; } tab[] = {
      {"ids", w.ids}, {"seq", w.seq}, {"idflag", w.idflag}, {"X5", w.X5}, {"sgate", w.sgate}, {"xhat", w.xhat}, {"rstd", w.rstd}, {"cnt", w.cnt},
      {"C", w.C}, {"T", w.T}, {"H1", w.Hd1}, {"A1", w.A1}, {"H2", w.Hd2}, {"A2", w.A2}, {"logit", w.logit}, {"prob", w.prob},
      {"bn", w.bn}, {"dlogit", w.dlogit}, {"dH2", w.dH2}, {"dH1", w.dH1}, {"dC", w.dC}, {"dT", w.dT}, {"dV", w.dV},
      {"dXitem", w.dXitem}, {"dXhist", w.dXhist}, {"dln", w.dln}, {"dy", w.dy}, {"row_off", w.row_off}, {"row_cnt", w.row_cnt},
      {"vals_out", w.vals_out}, {"keys_out", w.keys_out}};
  for (auto& t : tab)
    if (strcmp(t.n, name) == 0) return (size_t)((char*)t.p - base);
  return (size_t)-1;
}

// ---- branch parallelism inside the backward pass -------------------------------------------------
// The weight-gradient GEMMs, bias / LayerNorm / SENET / cate_emb gradient reductions are leaves of the dependency graph:
// they run on a library-owned side stream (fork = event on the caller's stream, join before the gradient norm), so the
// critical path is only the data-gradient chain.  Works under CUDA-graph capture (event fork/join is capturable); the
// stream and events are created on the first non-capturing call.
struct SideCtx {
  cudaStream_t s = nullptr;
  cudaEvent_t fork_ev[6], join_ev, mark_ev;
  int dev = -1;
  bool ok = false;
};
static SideCtx g_side;
static int g_use_side = 1;
static int g_ext_proj = 1;      // fbn_set_option("ext_proj", 0): keep the item_emb_d128 projection inside the gather kernel (A/B runs)

static bool side_ready(cudaStream_t main) {
  if (!g_use_side) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (g_side.ok && g_side.dev == dev) return true;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(main, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
  if (cudaStreamCreateWithFlags(&g_side.s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
  for (auto& e : g_side.fork_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&g_side.join_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&g_side.mark_ev, cudaEventDisableTiming);
  g_side.dev = dev; g_side.ok = true;
  return true;
}
// side stream waits for everything issued so far on `main`
static int side_fork(cudaStream_t main, int slot) {
  FBN_CHECK_CUDA(cudaEventRecord(g_side.fork_ev[slot], main));
  FBN_CHECK_CUDA(cudaStreamWaitEvent(g_side.s, g_side.fork_ev[slot], 0));
  return FBN_OK;
}
// a point on the side stream the caller's stream can wait for BEFORE the final join (everything issued on the side stream so far)
static int side_mark() { FBN_CHECK_CUDA(cudaEventRecord(g_side.mark_ev, g_side.s)); return FBN_OK; }
static int side_wait_mark(cudaStream_t main) { FBN_CHECK_CUDA(cudaStreamWaitEvent(main, g_side.mark_ev, 0)); return FBN_OK; }
static int side_join(cudaStream_t main) {
  FBN_CHECK_CUDA(cudaEventRecord(g_side.join_ev, g_side.s));
  FBN_CHECK_CUDA(cudaStreamWaitEvent(main, g_side.join_ev, 0));
  return FBN_OK;
}

// ---- bilinear transforms T = V_src * W_idx ---------------------------------------------------
static int bilinear_transform_fwd(const fbn_params_t* p, Workspace& w, cudaStream_t st) {
  const int type = p->bilinear_type;
  const int nT = type == FBN_BILINEAR_INTERACTION ? 10 : 4;
  GemmArgs g;
  g.M = w.B; g.N = D; g.K = D; g.lda = K1; g.ldb = D; g.ldc = nT * D; g.a_t = 0; g.b_t = 0;
  if (type == FBN_BILINEAR_ALL) {           // T_t = V_{t+2} W
    g.A = w.C + 2 * D; g.strideA = D; g.B = p->bil_w; g.strideB = 0; g.C = w.T; g.strideC = D; g.batch = 4;
    return tl_reg.run(g, w);
  }
  if (type == FBN_BILINEAR_EACH) {          // T_t = V_{t+1} W_{t+1}
    g.A = w.C + 1 * D; g.strideA = D; g.B = p->bil_w + 1 * D * D; g.strideB = D * D; g.C = w.T; g.strideC = D; g.batch = 4;
    return tl_reg.run(g, w);
  }
  int q0 = 0;                               // T_q = V_i W_(i,j), grouped by i
  for (int i = 1; i < NF - 1; ++i) {
    const int nj = NF - 1 - i;
    const int pidx = i * (2 * NF - i - 1) / 2;  // pair index of (i, i+1) in the full enumeration
    g.A = w.C + i * D; g.strideA = 0; g.B = p->bil_w + (long long)pidx * D * D; g.strideB = D * D;
    g.C = w.T + q0 * D; g.strideC = D; g.batch = nj;
    RC(tl_reg.run(g, w));
    q0 += nj;
  }
  return FBN_OK;
}

// Hadamard pairs -> the MLP input in GEMM operand format.  tf32x3 / bf16: the pair blocks join the field blocks the gather kernel
// already wrote into pk_C.  f16x3: the whole live row is written as fp16 hi|lo under one scale into pk16_C (pk_C keeps the tf32
// copy of the field blocks for the short-K bilinear GEMMs).
static int pairs_into_mlp_input(const fbn_params_t* p, Workspace& w, const PackDst& pkC, cudaStream_t st) {
  if (!tl_reg.mlp16()) return bilinear_pairs_fwd(p->bilinear_type, w.C, w.T, w.B, pkC, st);
  tl_reg.describe_mlp(w.C, w.B, K1, w.pk16_C);
  const Packed pk = packed_describe(w.pk16_C, w.B, K1, FBN_PREC_F16X3);
  return bilinear_pairs_mlp16(p->bilinear_type, w.C, w.T, w.B, pk.data, pk.lo_off, pk.scale, st);
}

// The item_emb_d128 projection (ref :106) runs on the tensor cores BEFORE the gather kernel when the batch carries the vectors:
// one short-K GEMM  Y[B,128] = item_mm[B,128] x mm_w^T + mm_b  (persistent tile loop), whose operand copy of item_mm is the one
// the mm_proj.0.weight gradient reads later.  The gather kernel then needs neither W in shared memory nor its SIMT projection loop
// (ncu: 30.9 M shared-memory wavefronts, 49 % SM throughput at 27 % DRAM).  With a resident mm_table the rows are only known inside
// the gather, which keeps its own projection.  tl_reg must be set up by the caller.
static int run_embed_fwd(const fbn_params_t* p, const fbn_batch_t* b, Workspace& w, int save, cudaStream_t st,
                         PackDst pkC = PackDst(), PackDst pkX = PackDst()) {
  const long long B = b->batch;
  EmbedFwdArgs e{};
  if (b->item_mm && g_ext_proj) {
    RC(tl_reg.pack(p->mm_w, D, D, w.pk_mmw));
    RC(tl_reg.pack(b->item_mm, B, D, w.pk_xmm));
    GemmArgs g;
    g.A = b->item_mm; g.lda = D; g.B = p->mm_w; g.ldb = D; g.b_t = 1; g.bias = p->mm_b; g.C = w.Ymm; g.ldc = D; g.M = B; g.N = D; g.K = D;
    RC(tl_reg.run(g, w));
    e.yproj = w.Ymm;
    pkX = PackDst();
  }
  e.item_emb = p->item_emb; e.cate_emb = p->cate_emb; e.mm_w = p->mm_w; e.mm_b = p->mm_b; e.ln_g = p->ln_g; e.ln_b = p->ln_b;
  e.se_w1 = p->se_w1; e.se_b1 = p->se_b1; e.se_w2 = p->se_w2; e.se_b2 = p->se_b2;
  e.item_id = b->item_id; e.likes = b->likes_level; e.views = b->views_level;
  e.seq = b->seq_len > 0 ? b->item_seq : nullptr;
  e.item_mm = b->item_mm; e.mm_table = b->mm_table; e.idx_dtype = b->idx_dtype; e.seq_dtype = b->seq_dtype;
  e.B = B; e.L = (int)b->seq_len; e.item_rows = p->item_rows; e.cate_rows = (int)p->cate_rows; e.save = save;
  e.ids = w.ids; e.seq32 = w.seq; e.idflag = w.idflag; e.X5 = w.X5; e.sgate = w.sgate; e.xhat = w.xhat; e.xmm = w.xmm; e.rstd = w.rstd; e.cnt = w.cnt;
  e.C = w.C;
  e.pkC = pkC; e.pkX = pkX;
  e.nshard = p->n_shards;
  e.se_r = p->se_hidden > 0 ? p->se_hidden : SE_R_DEFAULT;
  for (int r = 0; r < p->n_shards; ++r) e.shard[r] = p->shard[r];
  return launch_embed_senet_fwd(e, st);
}

// G1+S1 alone: gather + pooling + projection + SENET -> C[:, 128:768] (and the saved tensors if save != 0)
extern "C" int fbn_embed_forward(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int save, fbn_stream_t stream) {
  RC(check_common(p, b, ws, ws_bytes));
  Workspace w;
  carve_workspace(w, ws, b->batch, b->seq_len, ws_rows(p));
  tl_reg = PkReg();
  tl_reg.prec = p->precision; tl_reg.st = (cudaStream_t)stream;
  return run_embed_fwd(p, b, w, save, (cudaStream_t)stream);
}

extern "C" int fbn_forward(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int train, float dropout_p,
                           const uint8_t* keep_mask1, const uint8_t* keep_mask2, uint64_t seed, uint64_t offset,
                           const int32_t* step_counter_dev, float* prob_out,
                           fbn_stream_t stream) {
  RC(check_common(p, b, ws, ws_bytes));
  FBN_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, FBN_ERR_ARG, "dropout_p must be in [0,1)");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  carve_workspace(w, ws, b->batch, b->seq_len, ws_rows(p));
  const long long B = b->batch;

  tl_reg = PkReg();
  tl_reg.prec = p->precision; tl_reg.st = st;
  const unsigned long long fmask = 0x3Eull;                       // MLP-input blocks 1..5: the SENET-weighted fields
  const int nW = p->bilinear_type == FBN_BILINEAR_ALL ? 1 : (p->bilinear_type == FBN_BILINEAR_EACH ? NF - 1 : FBN_PAIRS);
  // the weights are converted on the side stream while the gather kernel runs
  const bool parf = tl_reg.on() && side_ready(st);
  // (the bilinear weight first: the transforms wait only for it -- `mark` -- while the MLP weights, which in f16x3 take an amax
  // and a split pass each, keep converting beside the transforms and the pair stage; the MLP-1 GEMM joins the side stream)
  if (parf) { RC(side_fork(st, 4)); tl_reg.st = g_side.s; }
  RC(tl_reg.pack(p->bil_w, (long long)nW * D, D, w.pk_bil));
  if (parf) RC(side_mark());
  RC(tl_reg.pack_mlp(p->w1, H1, K1, w.pk_w1, active_mask()));
  RC(tl_reg.pack_mlp(p->w2, H2, H1, w.pk_w2));
  tl_reg.st = st;

  // activations are converted to the operand format by the kernels that produce them (no separate pack pass)
  const PackDst pkC = tl_reg.dst(w.C, B, K1, w.pk_C);
  STAGE("fwd:start");
  RC(run_embed_fwd(p, b, w, 1, st, pkC, tl_reg.dst(b->item_mm ? b->item_mm : w.xmm, B, D, w.pk_xmm)));
  (void)fmask;
  if (parf) RC(side_wait_mark(st));
  STAGE("fwd:embed+senet (bilinear weight ready)");
  RC(bilinear_transform_fwd(p, w, st));
  STAGE("fwd:bilinear transforms");
  RC(pairs_into_mlp_input(p, w, pkC, st));
  if (parf) RC(side_join(st));
  STAGE("fwd:bilinear pairs (join MLP weight packing)");

  float* mean1 = w.bn; float* rstd1 = w.bn + H1; float* mean2 = w.bn + 2 * H1; float* rstd2 = w.bn + 2 * H1 + H2;
  GemmArgs g1;
  g1.A = w.C; g1.B = p->w1; g1.bias = p->b1; g1.C = w.Hd1; g1.M = B; g1.N = H1; g1.K = K1; g1.lda = K1; g1.ldb = K1; g1.ldc = H1;
  g1.b_t = 1; g1.kmask = active_mask();
  RC(tl_reg.run(g1, w));
  STAGE("fwd:mlp1 gemm");
  if (train) RC(bn_train_stats(w.Hd1, B, H1, w.partial, mean1, rstd1, p->bn1_mean, p->bn1_var, st));
  else RC(bn_eval_stats(p->bn1_mean, p->bn1_var, H1, mean1, rstd1, st));
  DropArgs d1; d1.p = train ? dropout_p : 0.f; d1.mask = keep_mask1; d1.seed = seed; d1.offset = offset; d1.stream = 1; d1.step_dev = step_counter_dev;
  RC(bn_act(w.Hd1, mean1, rstd1, p->bn1_g, p->bn1_b, B, H1, d1, w.A1, tl_reg.dst_mlp(w.A1, B, H1, w.pk_A1), st));
  RC(tl_reg.after_mlp(w.A1, B, H1, w.pk_A1));
  STAGE("fwd:bn1 stats+act");

  GemmArgs g2;
  g2.A = w.A1; g2.B = p->w2; g2.bias = p->b2; g2.C = w.Hd2; g2.M = B; g2.N = H2; g2.K = H1; g2.lda = H1; g2.ldb = H1; g2.ldc = H2;
  g2.b_t = 1;
  RC(tl_reg.run(g2, w));
  STAGE("fwd:mlp2 gemm");
  if (train) RC(bn_train_stats(w.Hd2, B, H2, w.partial, mean2, rstd2, p->bn2_mean, p->bn2_var, st));
  else RC(bn_eval_stats(p->bn2_mean, p->bn2_var, H2, mean2, rstd2, st));
  DropArgs d2; d2.p = train ? dropout_p : 0.f; d2.mask = keep_mask2; d2.seed = seed; d2.offset = offset; d2.stream = 2; d2.step_dev = step_counter_dev;
  RC(head_fwd(w.Hd2, mean2, rstd2, p->bn2_g, p->bn2_b, p->w3, p->b3, B, d2, w.A2, w.logit, w.prob, st));
  STAGE("fwd:bn2 stats+head");
  if (prob_out) FBN_CHECK_CUDA(cudaMemcpyAsync(prob_out, w.prob, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  return FBN_OK;
}

// split-K factor for a weight gradient on the tcgen05 path: with CTA-pair tiles the CTA count is 2 * clusters * splits;
// choose the split that wastes the least of the last wave of 148 SMs
namespace fbn { int tc_reserved_sms(); }
static int pick_splits_pair(long long M, long long N, long long K, unsigned long long nmask, size_t max_floats) {
  // tiles of the persistent CTA-pair kernel: 256 rows x 2 LIVE 128-column blocks, walked by num_sms / 2 clusters
  long long nlive = 0;
  for (long long n0 = 0; n0 < N; n0 += 128)
    if (nmask == ~0ull || ((nmask >> (n0 / 128)) & 1ull)) ++nlive;
  const long long clusters = cdiv(M, 256) * cdiv(nlive, 2);
  const long long kblocks = std::max<long long>(1, cdiv(K, 32));
  const long long slots = std::max(1, (num_sms() - tc_reserved_sms()) / 2);   // clusters that can run at once
  // cost in k-block times of a 256 x 256 tile (~1.3 us measured): waves x k-blocks per split, plus the fixed-order reduction of the
  // partials (1.6 us per split for the 512 x 1920 MLP-1 gradient, proportional to the output size).  At K = 65536 the second
  // term is noise and the split that fills whole waves wins (9 for MLP-1: 144 tiles = 2 waves of 74); at the reference's batch
  // 4096 it is what matters (9 splits: 2 x 8 + 11 = 27 units; 4 splits: 16 + 5 = 21).
  const double red = 1.25 * (double)(M * nlive * 128) / (512.0 * 1920.0);
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 32; ++s) {
    if ((size_t)s * M * N > max_floats || cdiv(kblocks, s) < 4) break;
    const double cost = (double)cdiv(clusters * s, slots) * (double)cdiv(cdiv(K, 64), s) + red * s;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

static EmbGradArgs make_emb_args(const fbn_params_t* p, const fbn_batch_t* b, Workspace& w, int32_t* row_touched) {
  EmbGradArgs eg{};
  eg.item_id = b->item_id; eg.idx_dtype = b->idx_dtype;
  eg.seq = (b->seq_len > 0 && b->item_seq) ? b->item_seq : nullptr; eg.seq_dtype = b->seq_dtype;
  eg.B = b->batch; eg.L = (int)b->seq_len; eg.rows = p->item_rows;
  eg.dXitem = w.dXitem; eg.dXhist = w.dXhist; eg.keys_in = w.keys_in; eg.keys_out = w.keys_out; eg.vals_in = w.vals_in;
  eg.vals_out = w.vals_out; eg.row_count = row_touched ? row_touched : w.row_cnt; eg.row_off = w.row_off; eg.cub_tmp = w.cub_tmp;
  eg.cub_bytes = w.cub_bytes; eg.sumsq_partial = w.partial_embsq;
  return eg;
}

// Occurrence index of the embedding backward (sort of the B + B*L row ids, per-row counts / offsets).  It depends on the
// batch ids only, so a host may run it on a second stream concurrently with fbn_forward and pass index_ready = 1 to
// fbn_backward (engine.TrainStep does).
extern "C" int fbn_embed_index(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int32_t* row_touched,
                               fbn_stream_t stream) {
  RC(check_common(p, b, ws, ws_bytes));
  Workspace w;
  carve_workspace(w, ws, b->batch, b->seq_len, ws_rows(p));
  EmbGradArgs eg = make_emb_args(p, b, w, row_touched);
  return emb_index(eg, (cudaStream_t)stream);
}

static int wgrad(const float* dOut, long long ldo, const float* In, long long ldi, long long B, long long M, long long N,
                 unsigned long long nmask, int precision, Workspace& w, float* out, cudaStream_t st, float* scratch) {
  // out[M,N] = dOut[B,M]^T * In[B,N]
  GemmArgs g;
  g.A = dOut; g.lda = ldo; g.a_t = 1; g.B = In; g.ldb = ldi; g.b_t = 0; g.M = M; g.N = N; g.K = B; g.ldc = N;
  g.splits = pick_splits(cdiv(M, 128) * cdiv(N, 128), B);
  if (precision != FBN_PREC_FP32 && M > 128 && N >= 256) g.splits = pick_splits_pair(M, N, B, nmask, w.partial_floats);
  g.nmask = nmask;
  FBN_REQUIRE((size_t)g.splits * M * N <= w.partial_floats, FBN_ERR_ARG, "internal: split-K scratch too small");
  g.C = scratch; g.strideSplit = M * N;
  RC(tl_reg.run_on(g, w, st));
  return reduce_splits(scratch, g.splits, M, N, M * N, nmask, out, st);
}

// Backward pass in up to three phases (fbn_backward = all of them, interleaved over two streams):
//   FBN_BWD_CHAIN : the dependency chain  head -> BN2 -> dgrad2 -> BN1 -> dgrad1 -> bilinear -> SENET / projection -> table rows
//   FBN_BWD_LEAF1 : mlp.0.weight / mlp.0.bias gradients (89 % of the dense gradient bytes)
//   FBN_BWD_LEAF2 : every other leaf (mlp.4, bilinear W, cate_emb, SENET, LayerNorm, mm_proj gradients)
// A data-parallel host runs CHAIN first, starts the all-reduce of the table gradient, runs LEAF1 while it is in flight, starts
// the all-reduce of that bucket, runs LEAF2 (engine.TrainStep, world > 1): the collectives overlap the weight-gradient GEMMs.
static int backward_impl(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int train, float dropout_p,
                         const float* dprob, const fbn_grads_t* g, const float* dense_grad_flat, int64_t dense_grad_n,
                         float* item_grad, int32_t* row_touched, int zero_fill, int index_ready, float* grad_sumsq,
                         fbn_stream_t stream, int phases) {
  RC(check_common(p, b, ws, ws_bytes));
  FBN_REQUIRE(g && grad_sumsq, FBN_ERR_ARG, "fbn_backward: null pointer");
  FBN_REQUIRE(phases > 0 && phases <= 7, FBN_ERR_ARG, "fbn_backward_phase: phases must be a non-empty subset of CHAIN | LEAF1 | LEAF2");
  const bool all = phases == 7, chain = (phases & FBN_BWD_CHAIN) != 0, leaf1 = (phases & FBN_BWD_LEAF1) != 0, leaf2 = (phases & FBN_BWD_LEAF2) != 0;
  FBN_REQUIRE(dprob || !chain, FBN_ERR_ARG, "fbn_backward: null dprob");
  FBN_REQUIRE(item_grad || p->n_shards > 0 || !chain, FBN_ERR_ARG, "fbn_backward: item_grad may only be NULL for a row-sharded table");
  FBN_REQUIRE(!(item_grad && p->n_shards > 0), FBN_ERR_ARG, "fbn_backward: a row-sharded table takes its gradient through fbn_shard_*");
  FBN_REQUIRE(aligned16(item_grad), FBN_ERR_ALIGN, "item_grad is not 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  carve_workspace(w, ws, b->batch, b->seq_len, ws_rows(p));
  const long long B = b->batch;
  const int prec = p->precision;
  const float scale = (train && dropout_p > 0.f) ? 1.0f / (1.0f - dropout_p) : 1.0f;
  // operands packed by fbn_forward (and by an earlier phase) are still valid: re-register them (no launch), pack the new ones
  // as they appear
  const int type = p->bilinear_type;
  const int nT = type == FBN_BILINEAR_INTERACTION ? 10 : 4;
  tl_reg = PkReg();
  tl_reg.prec = p->precision; tl_reg.st = st;
  {
    const int nWr = p->bilinear_type == FBN_BILINEAR_ALL ? 1 : (p->bilinear_type == FBN_BILINEAR_EACH ? NF - 1 : FBN_PAIRS);
    tl_reg.describe_mlp(p->w1, H1, K1, w.pk_w1);
    tl_reg.describe_mlp(p->w2, H2, H1, w.pk_w2);
    tl_reg.describe(p->bil_w, (long long)nWr * D, D, w.pk_bil);
    tl_reg.describe(w.C, B, K1, w.pk_C);
    if (tl_reg.mlp16()) tl_reg.describe_mlp(w.C, B, K1, w.pk16_C);
    tl_reg.describe_mlp(w.A1, B, H1, w.pk_A1);
    tl_reg.describe(b->item_mm ? b->item_mm : w.xmm, B, D, w.pk_xmm);
    if (!chain) {      // produced by the CHAIN phase of an earlier call
      tl_reg.describe_mlp(w.dH2, B, H2, w.pk_dH2);
      tl_reg.describe_mlp(w.dH1, B, H1, w.pk_dH1);
      tl_reg.describe(w.dT, B, (long long)nT * D, w.pk_dT);
      tl_reg.describe(w.dy, B, D, w.pk_dy);
    }
  }
  float* mean1 = w.bn; float* rstd1 = w.bn + H1; float* mean2 = w.bn + 2 * H1; float* rstd2 = w.bn + 2 * H1 + H2;
  const unsigned long long amask = active_mask();
  // a leaf requested TOGETHER with the chain runs on the library's side stream, forked where its inputs become ready (so it
  // fills in beside the data-gradient chain); a leaf requested without the chain runs on the caller's stream
  const bool par = chain && (leaf1 || leaf2) && side_ready(st);
  cudaStream_t ls = par ? g_side.s : st;
  float* lp = par ? w.partial_side : w.partial;
  const int eb = embed_bwd_blocks(B);

  auto leaf_layer2 = [&]() -> int {
    RC(colsum(w.dH2, B, H2, lp, g->b2, ls));
    return wgrad(w.dH2, H2, w.A1, H1, B, H2, H1, ~0ull, prec, w, g->w2, ls, lp);
  };
  auto leaf_layer1 = [&]() -> int {
    RC(colsum(w.dH1, B, H1, lp, g->b1, ls));
    return wgrad(w.dH1, H1, w.C, K1, B, H1, K1, amask, prec, w, g->w1, ls, lp);
  };
  auto leaf_bilinear = [&]() -> int {
    // dW[idx] = sum_t V_src^T dT_t  (split-K over the batch, fixed-order reduction)
    GemmArgs d;
    d.a_t = 1; d.lda = K1; d.b_t = 0; d.ldb = nT * D; d.M = D; d.N = D; d.K = B; d.ldc = D;
    const int S = pick_splits(nT, B);
    d.splits = S; d.strideSplit = (long long)D * D; d.strideC = (long long)S * D * D; d.C = lp;
    FBN_REQUIRE((size_t)nT * S * D * D <= w.partial_floats, FBN_ERR_ARG, "internal: bilinear scratch too small");
    if (type == FBN_BILINEAR_ALL) {
      d.A = w.C + 2 * D; d.strideA = D; d.B = w.dT; d.strideB = D; d.batch = 4;
      RC(tl_reg.run_on(d, w, ls));
      RC(reduce_splits(lp, 4 * S, D, D, (long long)D * D, ~0ull, g->bil_w, ls));
    } else if (type == FBN_BILINEAR_EACH) {
      d.A = w.C + 1 * D; d.strideA = D; d.B = w.dT; d.strideB = D; d.batch = 4;
      RC(tl_reg.run_on(d, w, ls));
      FBN_CHECK_CUDA(cudaMemsetAsync(g->bil_w, 0, sizeof(float) * D * D, ls));  // W_0 multiplies the zero field
      for (int t = 0; t < 4; ++t)
        RC(reduce_splits(lp + (long long)t * S * D * D, S, D, D, (long long)D * D, ~0ull, g->bil_w + (long long)(t + 1) * D * D, ls));
    } else {
      FBN_CHECK_CUDA(cudaMemsetAsync(g->bil_w, 0, sizeof(float) * (NF - 1) * D * D, ls));  // pairs (0,j)
      int q = 0;
      for (int i = 1; i < NF - 1; ++i) {
        const int nj = NF - 1 - i;
        d.A = w.C + i * D; d.strideA = 0; d.B = w.dT + q * D; d.strideB = D; d.batch = nj; d.C = lp + (long long)q * S * D * D;
        RC(tl_reg.run_on(d, w, ls));
        q += nj;
      }
      for (int t = 0; t < 10; ++t)
        RC(reduce_splits(lp + (long long)t * S * D * D, S, D, D, (long long)D * D, ~0ull,
                         g->bil_w + (long long)(NF - 1 + t) * D * D, ls));
    }
    return FBN_OK;
  };
  auto leaf_embed = [&]() -> int {
    RC(launch_reduce_partials(w.partial_cate, g->cate_emb, eb, p->cate_rows * D, 0, ls));
    RC(launch_senet_param_grads(w.sestat, B, p->se_hidden > 0 ? p->se_hidden : SE_R_DEFAULT, lp, g->se_w1, g->se_b1, g->se_w2, g->se_b2, ls));
    RC(colprod2(w.dln, w.xhat, B, D, lp, g->ln_g, g->ln_b, ls));
    RC(colsum(w.dy, B, D, lp, g->mm_b, ls));
    return wgrad(w.dy, D, b->item_mm ? b->item_mm : w.xmm, D, B, D, D, ~0ull, prec, w, g->mm_w, ls, lp);
  };

  if (chain) {
    // ---- head + layer 2 ----
    STAGE("bwd:start");
    RC(head_bwd_stats(dprob, w.prob, w.A2, w.Hd2, mean2, rstd2, p->w3, B, scale, w.partial, w.dlogit, g->bn2_g, g->bn2_b, g->w3, g->b3,
                      st));
    RC(bn_bwd_apply(nullptr, w.dlogit, p->w3, w.A2, w.Hd2, mean2, rstd2, p->bn2_g, g->bn2_g, g->bn2_b, B, H2, scale, train, w.dH2,
                     tl_reg.dst_mlp(w.dH2, B, H2, w.pk_dH2), st));
    RC(tl_reg.after_mlp(w.dH2, B, H2, w.pk_dH2));
    STAGE("bwd:head + bn2");
    if (par && leaf2) RC(side_fork(st, 0));
    if (leaf2) RC(leaf_layer2());
    {
      GemmArgs d;  // dA1 = dH2 * w2
      d.A = w.dH2; d.lda = H2; d.B = p->w2; d.ldb = H1; d.b_t = 0; d.C = w.dH1; d.ldc = H1; d.M = B; d.N = H1; d.K = H2;
      RC(tl_reg.run(d, w));
    }
    STAGE("bwd:mlp2 dgrad (+side: wgrad2)");
    // ---- layer 1 ----
    RC(bn_bwd_stats(w.dH1, w.A1, w.Hd1, mean1, rstd1, B, H1, scale, w.partial, g->bn1_g, g->bn1_b, st));
    RC(bn_bwd_apply(w.dH1, nullptr, nullptr, w.A1, w.Hd1, mean1, rstd1, p->bn1_g, g->bn1_g, g->bn1_b, B, H1, scale, train, w.dH1,
                     tl_reg.dst_mlp(w.dH1, B, H1, w.pk_dH1), st));
    RC(tl_reg.after_mlp(w.dH1, B, H1, w.pk_dH1));
    STAGE("bwd:bn1");
    if (par && leaf1) RC(side_fork(st, 1));
    if (leaf1) RC(leaf_layer1());
    {
      GemmArgs d;  // dC = dH1 * w1 (only the blocks that feed something)
      d.A = w.dH1; d.lda = H1; d.B = p->w1; d.ldb = K1; d.b_t = 0; d.C = w.dC; d.ldc = K1; d.M = B; d.N = K1; d.K = H1; d.nmask = amask;
      RC(tl_reg.run(d, w));
    }
    STAGE("bwd:mlp1 dgrad (+side: wgrad1)");
    // ---- bilinear ----
    RC(bilinear_pairs_bwd(type, w.C, w.T, w.dC, B, w.dT, w.dV, tl_reg.dst(w.dT, B, (long long)nT * D, w.pk_dT), st));
    STAGE("bwd:bilinear pairs");
    {
      GemmArgs d;  // dV[src] += dT_t * W^T
      d.M = B; d.N = D; d.K = D; d.lda = nT * D; d.ldb = D; d.b_t = 1; d.ldc = NA * D; d.accumulate = 1;
      if (type == FBN_BILINEAR_ALL) {
        d.A = w.dT; d.strideA = D; d.B = p->bil_w; d.strideB = 0; d.C = w.dV + 1 * D; d.strideC = D; d.batch = 4;
        RC(tl_reg.run(d, w));
      } else if (type == FBN_BILINEAR_EACH) {
        d.A = w.dT; d.strideA = D; d.B = p->bil_w + D * D; d.strideB = D * D; d.C = w.dV; d.strideC = D; d.batch = 4;
        RC(tl_reg.run(d, w));
      } else {
        int q = 0;
        for (int i = 1; i < NF - 1; ++i)
          for (int j = i + 1; j < NF; ++j, ++q) {
            const int pidx = i * (2 * NF - i - 1) / 2 + (j - i - 1);
            d.A = w.dT + q * D; d.B = p->bil_w + (long long)pidx * D * D; d.C = w.dV + (i - 1) * D; d.batch = 1;
            RC(tl_reg.run(d, w));
          }
      }
    }
    STAGE("bwd:bilinear dgrad");
    if (par && leaf2) RC(side_fork(st, 2));
    if (leaf2) RC(leaf_bilinear());
    // ---- SENET + field stack + projection ----
    EmbedBwdArgs e{};
    e.dV = w.dV; e.X5 = w.X5; e.sgate = w.sgate; e.xhat = w.xhat; e.rstd = w.rstd; e.cnt = w.cnt; e.ids = w.ids;
    e.se_w1 = p->se_w1; e.se_b1 = p->se_b1; e.se_w2 = p->se_w2; e.ln_g = p->ln_g; e.B = B; e.cate_rows = (int)p->cate_rows;
    e.se_r = p->se_hidden > 0 ? p->se_hidden : SE_R_DEFAULT;
    e.dXitem = w.dXitem; e.dXhist = w.dXhist; e.dln = w.dln; e.dy = w.dy; e.sestat = w.sestat; e.cate_partial = w.partial_cate;
    e.pkdy = tl_reg.dst(w.dy, B, D, w.pk_dy);
    FBN_REQUIRE((size_t)eb * p->cate_rows * D <= (size_t)2 * 148 * 4 * MAX_CATE * D, FBN_ERR_ARG, "internal: cate scratch too small");
    RC(launch_embed_senet_bwd(e, eb, st));
    STAGE("bwd:embed+senet");
    if (par && leaf2) RC(side_fork(st, 3));
    if (leaf2) RC(leaf_embed());
    // ---- embedding table rows ----
    if (item_grad) {
      EmbGradArgs eg = make_emb_args(p, b, w, row_touched);
      eg.grad = item_grad; eg.zero_fill = zero_fill; eg.sumsq_out = grad_sumsq + 1;
      if (!index_ready) RC(emb_index(eg, st));
      RC(emb_rows(eg, st));
    }
    STAGE("bwd:table rows");
    if (par) RC(side_join(st));     // every dense gradient is complete from here on
    STAGE("bwd:join side stream (leaf gradients)");
  }
  if (!chain) {
    if (leaf1) RC(leaf_layer1());
    if (leaf2) {
      RC(leaf_layer2());
      RC(leaf_bilinear());
      RC(leaf_embed());
    }
  }
  if (!all) return FBN_OK;      // the host computes the gradient norms after its collectives
  if (dense_grad_flat) {
    FBN_REQUIRE(aligned16(dense_grad_flat), FBN_ERR_ALIGN, "dense_grad_flat is not 16-byte aligned");
    RC(sumsq(dense_grad_flat, dense_grad_n, w.partial, grad_sumsq, st));
  }
  STAGE("bwd:dense sumsq");
  return FBN_OK;
}

extern "C" int fbn_backward(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int train, float dropout_p,
                            const float* dprob, const fbn_grads_t* g, const float* dense_grad_flat, int64_t dense_grad_n,
                            float* item_grad, int32_t* row_touched, int zero_fill, int index_ready, float* grad_sumsq,
                            fbn_stream_t stream) {
  return backward_impl(p, b, ws, ws_bytes, train, dropout_p, dprob, g, dense_grad_flat, dense_grad_n, item_grad, row_touched, zero_fill,
                       index_ready, grad_sumsq, stream, 7);
}

extern "C" int fbn_backward_phase(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int train, float dropout_p,
                                  const float* dprob, const fbn_grads_t* g, float* item_grad, int32_t* row_touched, int zero_fill,
                                  int index_ready, float* grad_sumsq, int phases, fbn_stream_t stream) {
  return backward_impl(p, b, ws, ws_bytes, train, dropout_p, dprob, g, nullptr, 0, item_grad, row_touched, zero_fill, index_ready,
                       grad_sumsq, stream, phases);
}

// ---- MLP tower for an arbitrary input width (F-field model, general.py) ------------------------------------------------------
namespace fbn {
struct TowerWs {
  float *Hd1, *A1, *Hd2, *A2, *logit, *prob, *bn, *dlogit, *dH2, *dH1, *partial, *partial_side;
  size_t partial_floats;
  void *pk_C, *pk_A1, *pk_dH2, *pk_dH1, *pk_w1, *pk_w2;
  size_t total_bytes;
};

static void carve_tower(TowerWs& w, void* base, int64_t B, int64_t k1) {
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? p + off : nullptr;
    off += align_up(bytes);
    return r;
  };
  const size_t f = sizeof(float);
  const int64_t Bp = std::max<int64_t>(B, 1);
  w.Hd1 = (float*)take(Bp * H1 * f); w.A1 = (float*)take(Bp * H1 * f);
  w.Hd2 = (float*)take(Bp * H2 * f); w.A2 = (float*)take(Bp * H2 * f);
  w.logit = (float*)take(Bp * f); w.prob = (float*)take(Bp * f);
  w.bn = (float*)take((2 * H1 + 2 * H2) * f);
  w.dlogit = (float*)take(Bp * f);
  w.dH2 = (float*)take(Bp * H2 * f); w.dH1 = (float*)take(Bp * H1 * f);
  size_t pf = std::max<size_t>((size_t)4 * H1 * k1, (size_t)32 * H2 * H1);   // split-K partials of the two weight gradients
  pf = std::max<size_t>(pf, (size_t)16 * 148 * 3 * H1);
  w.partial_floats = pf;
  w.partial = (float*)take(pf * f);
  w.partial_side = (float*)take((size_t)16 * 148 * 3 * H1 * f);
  auto pkb = [](long long r, long long c) { return std::max(packed_bytes(r, c, FBN_PREC_TF32X3), packed_bytes(r, c, FBN_PREC_F16X3)); };
  w.pk_C = take(pkb(Bp, k1));
  w.pk_A1 = take(pkb(Bp, H1));
  w.pk_dH2 = take(pkb(Bp, H2));
  w.pk_dH1 = take(pkb(Bp, H1));
  w.pk_w1 = take(pkb(H1, k1));
  w.pk_w2 = take(pkb(H2, H1));
  w.total_bytes = off;
}

static int tower_check(const fbn_params_t* p, const float* c, int64_t batch, int64_t k1, void* ws, size_t ws_bytes) {
  FBN_REQUIRE(p && c && ws, FBN_ERR_ARG, "fbn_tower: null pointer");
  FBN_REQUIRE(batch >= 1 && k1 >= 128 && k1 % 128 == 0, FBN_ERR_SHAPE, "fbn_tower: the input width must be a multiple of 128");
  FBN_REQUIRE((p->precision >= FBN_PREC_FP32 && p->precision <= FBN_PREC_BF16) || p->precision == FBN_PREC_F16X3, FBN_ERR_ARG, "bad precision");
  const void* ptrs[] = {c, ws, p->w1, p->b1, p->bn1_g, p->bn1_b, p->bn1_mean, p->bn1_var, p->w2, p->b2, p->bn2_g, p->bn2_b, p->bn2_mean,
                        p->bn2_var, p->w3};
  for (const void* q : ptrs) FBN_REQUIRE(q && aligned16(q), FBN_ERR_ALIGN, "fbn_tower: a tensor pointer is missing or not 16-byte aligned");
  FBN_REQUIRE(p->b3 != nullptr, FBN_ERR_ARG, "fbn_tower: null b3");
  TowerWs w;
  carve_tower(w, nullptr, batch, k1);
  FBN_REQUIRE(ws_bytes >= w.total_bytes, FBN_ERR_ARG, "fbn_tower: workspace too small: %zu < %zu", ws_bytes, w.total_bytes);
  return FBN_OK;
}
}  // namespace fbn

extern "C" size_t fbn_tower_workspace_bytes(int64_t batch, int64_t k1) {
  TowerWs w;
  carve_tower(w, nullptr, batch, k1);
  return w.total_bytes;
}

extern "C" size_t fbn_tower_workspace_offset(int64_t batch, int64_t k1, const char* name) {
  TowerWs w;
  char* base = reinterpret_cast<char*>(uintptr_t(4096));
  carve_tower(w, base, batch, k1);
  struct { const char* n; void* p; } tab[] = {{"H1", w.Hd1}, {"A1", w.A1}, {"H2", w.Hd2}, {"A2", w.A2}, {"logit", w.logit}, {"prob", w.prob},
                                              {"dH1", w.dH1}, {"dH2", w.dH2}, {"dlogit", w.dlogit}};
  for (auto& t : tab)
    if (strcmp(t.n, name) == 0) return (size_t)((char*)t.p - base);
  return (size_t)-1;
}

extern "C" int fbn_tower_forward(const fbn_params_t* p, const float* c, int64_t batch, int64_t k1, void* ws, size_t ws_bytes, int train,
                                 float dropout_p, const uint8_t* keep_mask1, const uint8_t* keep_mask2, uint64_t seed, uint64_t offset,
                                 const int32_t* step_counter_dev, float* prob_out, float* logit_out, fbn_stream_t stream) {
  RC(tower_check(p, c, batch, k1, ws, ws_bytes));
  FBN_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, FBN_ERR_ARG, "dropout_p must be in [0,1)");
  cudaStream_t st = (cudaStream_t)stream;
  TowerWs w;
  carve_tower(w, ws, batch, k1);
  Workspace shell{};                       // the GEMM helpers only look at the operand scratch (unused: everything is pre-packed)
  const long long B = batch;
  tl_reg = PkReg();
  tl_reg.prec = p->precision; tl_reg.st = st;
  RC(tl_reg.pack_mlp(p->w1, H1, k1, w.pk_w1));     // every GEMM of the tower is a long-K MLP GEMM: all operands in the MLP format
  RC(tl_reg.pack_mlp(p->w2, H2, H1, w.pk_w2));
  RC(tl_reg.pack_mlp(c, B, k1, w.pk_C));
  float* mean1 = w.bn; float* rstd1 = w.bn + H1; float* mean2 = w.bn + 2 * H1; float* rstd2 = w.bn + 2 * H1 + H2;
  GemmArgs g1;
  g1.A = c; g1.B = p->w1; g1.bias = p->b1; g1.C = w.Hd1; g1.M = B; g1.N = H1; g1.K = k1; g1.lda = k1; g1.ldb = k1; g1.ldc = H1; g1.b_t = 1;
  RC(tl_reg.run(g1, shell));
  if (train) RC(bn_train_stats(w.Hd1, B, H1, w.partial, mean1, rstd1, p->bn1_mean, p->bn1_var, st));
  else RC(bn_eval_stats(p->bn1_mean, p->bn1_var, H1, mean1, rstd1, st));
  DropArgs d1; d1.p = train ? dropout_p : 0.f; d1.mask = keep_mask1; d1.seed = seed; d1.offset = offset; d1.stream = 1; d1.step_dev = step_counter_dev;
  RC(bn_act(w.Hd1, mean1, rstd1, p->bn1_g, p->bn1_b, B, H1, d1, w.A1, tl_reg.dst_mlp(w.A1, B, H1, w.pk_A1), st));
  RC(tl_reg.after_mlp(w.A1, B, H1, w.pk_A1));
  GemmArgs g2;
  g2.A = w.A1; g2.B = p->w2; g2.bias = p->b2; g2.C = w.Hd2; g2.M = B; g2.N = H2; g2.K = H1; g2.lda = H1; g2.ldb = H1; g2.ldc = H2; g2.b_t = 1;
  RC(tl_reg.run(g2, shell));
  if (train) RC(bn_train_stats(w.Hd2, B, H2, w.partial, mean2, rstd2, p->bn2_mean, p->bn2_var, st));
  else RC(bn_eval_stats(p->bn2_mean, p->bn2_var, H2, mean2, rstd2, st));
  DropArgs d2; d2.p = train ? dropout_p : 0.f; d2.mask = keep_mask2; d2.seed = seed; d2.offset = offset; d2.stream = 2; d2.step_dev = step_counter_dev;
  RC(head_fwd(w.Hd2, mean2, rstd2, p->bn2_g, p->bn2_b, p->w3, p->b3, B, d2, w.A2, w.logit, w.prob, st));
  if (prob_out) FBN_CHECK_CUDA(cudaMemcpyAsync(prob_out, w.prob, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  if (logit_out) FBN_CHECK_CUDA(cudaMemcpyAsync(logit_out, w.logit, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  return FBN_OK;
}

extern "C" int fbn_tower_backward(const fbn_params_t* p, const float* c, int64_t batch, int64_t k1, void* ws, size_t ws_bytes, int train,
                                  float dropout_p, const float* dprob, const fbn_grads_t* g, float* dc, fbn_stream_t stream) {
  RC(tower_check(p, c, batch, k1, ws, ws_bytes));
  FBN_REQUIRE(dprob && g && dc && aligned16(dc), FBN_ERR_ARG, "fbn_tower_backward: null / unaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  TowerWs w;
  carve_tower(w, ws, batch, k1);
  Workspace shell{};
  shell.partial_floats = w.partial_floats;
  const long long B = batch;
  const int prec = p->precision;
  const float scale = (train && dropout_p > 0.f) ? 1.0f / (1.0f - dropout_p) : 1.0f;
  tl_reg = PkReg();
  tl_reg.prec = p->precision; tl_reg.st = st;
  tl_reg.describe_mlp(p->w1, H1, k1, w.pk_w1);
  tl_reg.describe_mlp(p->w2, H2, H1, w.pk_w2);
  tl_reg.describe_mlp(c, B, k1, w.pk_C);
  tl_reg.describe_mlp(w.A1, B, H1, w.pk_A1);
  float* mean1 = w.bn; float* rstd1 = w.bn + H1; float* mean2 = w.bn + 2 * H1; float* rstd2 = w.bn + 2 * H1 + H2;
  RC(head_bwd_stats(dprob, w.prob, w.A2, w.Hd2, mean2, rstd2, p->w3, B, scale, w.partial, w.dlogit, g->bn2_g, g->bn2_b, g->w3, g->b3, st));
  RC(bn_bwd_apply(nullptr, w.dlogit, p->w3, w.A2, w.Hd2, mean2, rstd2, p->bn2_g, g->bn2_g, g->bn2_b, B, H2, scale, train, w.dH2,
                   tl_reg.dst_mlp(w.dH2, B, H2, w.pk_dH2), st));
  RC(tl_reg.after_mlp(w.dH2, B, H2, w.pk_dH2));
  RC(colsum(w.dH2, B, H2, w.partial, g->b2, st));
  RC(wgrad(w.dH2, H2, w.A1, H1, B, H2, H1, ~0ull, prec, shell, g->w2, st, w.partial));
  {
    GemmArgs d;  // dA1 = dH2 * w2
    d.A = w.dH2; d.lda = H2; d.B = p->w2; d.ldb = H1; d.b_t = 0; d.C = w.dH1; d.ldc = H1; d.M = B; d.N = H1; d.K = H2;
    RC(tl_reg.run(d, shell));
  }
  RC(bn_bwd_stats(w.dH1, w.A1, w.Hd1, mean1, rstd1, B, H1, scale, w.partial, g->bn1_g, g->bn1_b, st));
  RC(bn_bwd_apply(w.dH1, nullptr, nullptr, w.A1, w.Hd1, mean1, rstd1, p->bn1_g, g->bn1_g, g->bn1_b, B, H1, scale, train, w.dH1,
                   tl_reg.dst_mlp(w.dH1, B, H1, w.pk_dH1), st));
  RC(tl_reg.after_mlp(w.dH1, B, H1, w.pk_dH1));
  RC(colsum(w.dH1, B, H1, w.partial, g->b1, st));
  RC(wgrad(w.dH1, H1, c, k1, B, H1, k1, ~0ull, prec, shell, g->w1, st, w.partial));
  {
    GemmArgs d;  // dC = dH1 * w1
    d.A = w.dH1; d.lda = H1; d.B = p->w1; d.ldb = k1; d.b_t = 0; d.C = dc; d.ldc = k1; d.M = B; d.N = k1; d.K = H1;
    RC(tl_reg.run(d, shell));
  }
  return FBN_OK;
}

// Per-stage device times of the calls made since the last report (see stage_mark): "name<TAB>ms" lines.  Synchronises.
extern "C" int fbn_stage_report(char* buf, size_t buf_bytes) {
  FBN_REQUIRE(buf && buf_bytes > 0, FBN_ERR_ARG, "fbn_stage_report: null buffer");
  FBN_CHECK_CUDA(cudaDeviceSynchronize());
  std::string out;
  for (size_t i = 1; i < g_stage.ev.size(); ++i) {
    float ms = 0.f;
    if (g_stage.ev[i].first.find(":start") == std::string::npos) {
      FBN_CHECK_CUDA(cudaEventElapsedTime(&ms, g_stage.ev[i - 1].second, g_stage.ev[i].second));
      char line[160];
      snprintf(line, sizeof(line), "%s\t%.4f\n", g_stage.ev[i].first.c_str(), ms);
      out += line;
    }
  }
  for (auto& e : g_stage.ev) cudaEventDestroy(e.second);
  g_stage.ev.clear();
  snprintf(buf, buf_bytes, "%s", out.c_str());
  return FBN_OK;
}

// Benchmark helper: mean CUDA-event duration of ONE stage of the forward pass exactly as fbn_forward runs it (same operands,
// same strides, same launch configuration), after a complete fbn_forward on the same workspace.  stage: "bil_gemm"
// (bilinear transforms), "bil_pairs" (Hadamard pairs into C), "embed" (gather + SENET), "mlp1" (MLP-1 forward GEMM).
extern "C" int fbn_time_stage(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, const char* stage, void* flush,
                              size_t flush_bytes, int iters, float* ms_out, fbn_stream_t stream) {
  RC(check_common(p, b, ws, ws_bytes));
  FBN_REQUIRE(stage && ms_out && iters > 0, FBN_ERR_ARG, "fbn_time_stage: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  carve_workspace(w, ws, b->batch, b->seq_len, ws_rows(p));
  const long long B = b->batch;
  tl_reg = PkReg();
  tl_reg.prec = p->precision; tl_reg.st = st;
  const int nW = p->bilinear_type == FBN_BILINEAR_ALL ? 1 : (p->bilinear_type == FBN_BILINEAR_EACH ? NF - 1 : FBN_PAIRS);
  tl_reg.describe_mlp(p->w1, H1, K1, w.pk_w1);
  tl_reg.describe_mlp(p->w2, H2, H1, w.pk_w2);
  tl_reg.describe(p->bil_w, (long long)nW * D, D, w.pk_bil);
  const PackDst pkC = tl_reg.dst(w.C, B, K1, w.pk_C);
  if (tl_reg.mlp16()) tl_reg.describe_mlp(w.C, B, K1, w.pk16_C);
  const PackDst pkX = tl_reg.dst(b->item_mm ? b->item_mm : w.xmm, B, D, w.pk_xmm);
  tl_reg.describe_mlp(w.A1, B, H1, w.pk_A1);
  tl_reg.describe_mlp(w.dH1, B, H1, w.pk_dH1);      // valid after an fbn_backward on this workspace (mlp1_dgrad / mlp1_wgrad)
  const std::string s(stage);
  cudaEvent_t e0, e1;
  FBN_CHECK_CUDA(cudaEventCreate(&e0));
  FBN_CHECK_CUDA(cudaEventCreate(&e1));
  float total = 0.f;
  for (int it = 0; it < iters + 1; ++it) {
    if (flush) FBN_CHECK_CUDA(cudaMemsetAsync(flush, it, flush_bytes, st));
    FBN_CHECK_CUDA(cudaEventRecord(e0, st));
    if (s == "bil_gemm") RC(bilinear_transform_fwd(p, w, st));
    else if (s == "bil_pairs") RC(pairs_into_mlp_input(p, w, pkC, st));
    else if (s == "embed") RC(run_embed_fwd(p, b, w, 1, st, pkC, pkX));
    else if (s == "mlp1") {
      GemmArgs g1;
      g1.A = w.C; g1.B = p->w1; g1.bias = p->b1; g1.C = w.Hd1; g1.M = B; g1.N = H1; g1.K = K1; g1.lda = K1; g1.ldb = K1; g1.ldc = H1;
      g1.b_t = 1; g1.kmask = active_mask();
      RC(tl_reg.run(g1, w));
    } else if (s == "mlp1_dgrad") {      // dC = dH1 * w1, live column blocks only (as in fbn_backward)
      GemmArgs d;
      d.A = w.dH1; d.lda = H1; d.B = p->w1; d.ldb = K1; d.b_t = 0; d.C = w.dC; d.ldc = K1; d.M = B; d.N = K1; d.K = H1; d.nmask = active_mask();
      RC(tl_reg.run(d, w));
    } else if (s == "mlp1_wgrad") {      // dw1 = dH1^T * C, split-K + fixed-order reduce (as in fbn_backward; result goes to scratch)
      RC(wgrad(w.dH1, H1, w.C, K1, B, H1, K1, active_mask(), p->precision, w, w.partial_side, st, w.partial));
    } else {
      FBN_REQUIRE(false, FBN_ERR_ARG, "fbn_time_stage: unknown stage '%s'", stage);
    }
    FBN_CHECK_CUDA(cudaEventRecord(e1, st));
    FBN_CHECK_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    FBN_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0) total += ms;      // first iteration = warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = total / iters;
  return FBN_OK;
}

// ---- BCELoss ---------------------------------------------------------------------------------
namespace fbn {
__global__ void bce_kernel(const float* __restrict__ prob, const float* __restrict__ y, long long B, float scale, float* loss_out,
                           float* __restrict__ dprob) {
  // single block: B <= a few 100k, fixed-order double accumulation
  __shared__ double s[256];
  double t = 0.0;
  const float invB = 1.0f / (float)B;
  for (long long i = threadIdx.x; i < B; i += 256) {
    const float p = prob[i], yy = y[i];
    const float lp = fmaxf(logf(p), -100.0f), l1p = fmaxf(log1pf(-p), -100.0f);
    t += (double)(-(yy * lp + (1.0f - yy) * l1p));
    if (dprob) dprob[i] = scale * ((p - yy) / fmaxf((1.0f - p) * p, 1e-12f)) * invB;
  }
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss_out) loss_out[0] = (float)(s[0] / (double)B);
}
}  // namespace fbn

namespace fbn {
// multi-block BCELoss: every block handles one contiguous chunk (dprob + a double partial of the loss); the block that arrives
// last adds the partials in block order (fixed order -> run-to-run deterministic) and re-arms the counter for the next launch.
// (The single-block kernel above costs 148 us at B = 65536 -- on the critical path between forward and backward.)
constexpr int BCE_MAX_BLOCKS = 512;
__global__ void __launch_bounds__(256) bce_multi_kernel(const float* __restrict__ prob, const float* __restrict__ y, long long B, long long per,
                                                        float scale, float* loss_out, float* __restrict__ dprob, double* partial,
                                                        unsigned int* counter) {
  __shared__ double s[256];
  __shared__ bool last;
  const long long b0 = (long long)blockIdx.x * per, b1 = min(B, b0 + per);
  const float invB = 1.0f / (float)B;
  double t = 0.0;
  for (long long i = b0 + threadIdx.x; i < b1; i += 256) {
    const float p = prob[i], yy = y[i];
    const float lp = fmaxf(logf(p), -100.0f), l1p = fmaxf(log1pf(-p), -100.0f);
    t += (double)(-(yy * lp + (1.0f - yy) * l1p));
    if (dprob) dprob[i] = scale * ((p - yy) / fmaxf((1.0f - p) * p, 1e-12f)) * invB;
  }
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s[0];
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double u = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) u += ((volatile double*)partial)[i];
  s[threadIdx.x] = u;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (loss_out) loss_out[0] = (float)(s[0] / (double)B);
    *counter = 0;
  }
}
}  // namespace fbn

extern "C" int fbn_bce_loss(const float* prob, const float* labels, int64_t batch, float loss_scale, float* loss_out, float* dprob_out,
                            fbn_stream_t stream) {
  FBN_REQUIRE(prob && labels && batch >= 1, FBN_ERR_ARG, "fbn_bce_loss: bad arguments");
  bce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(prob, labels, batch, loss_scale, loss_out, dprob_out);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" size_t fbn_bce_scratch_bytes(void) { return (size_t)BCE_MAX_BLOCKS * sizeof(double) + 64; }

extern "C" int fbn_bce_loss_ws(const float* prob, const float* labels, int64_t batch, float loss_scale, float* loss_out, float* dprob_out,
                               void* scratch, size_t scratch_bytes, fbn_stream_t stream) {
  FBN_REQUIRE(prob && labels && batch >= 1 && scratch, FBN_ERR_ARG, "fbn_bce_loss_ws: bad arguments");
  FBN_REQUIRE(scratch_bytes >= fbn_bce_scratch_bytes() && (reinterpret_cast<uintptr_t>(scratch) & 7u) == 0, FBN_ERR_ARG,
              "fbn_bce_loss_ws: scratch too small / unaligned");
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(batch, 1024), std::min(BCE_MAX_BLOCKS, 2 * num_sms())));
  const long long per = cdiv(batch, blocks);
  double* partial = static_cast<double*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + BCE_MAX_BLOCKS);
  bce_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(prob, labels, batch, per, loss_scale, loss_out, dprob_out, partial, counter);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

extern "C" int fbn_gemm(const float* A, const float* Bm, const float* bias, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                        int64_t ldb, int64_t ldc, int a_t, int b_t, int precision, void* scratch, size_t scratch_bytes,
                        fbn_stream_t stream) {
  FBN_REQUIRE(A && Bm && C, FBN_ERR_ARG, "fbn_gemm: null pointer");
  FBN_REQUIRE(aligned16(A) && aligned16(Bm) && aligned16(C) && aligned16(bias), FBN_ERR_ALIGN, "fbn_gemm: unaligned pointer");
  GemmArgs g;
  g.A = A; g.B = Bm; g.bias = bias; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.a_t = a_t; g.b_t = b_t;
  return gemm(g, precision, scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" size_t fbn_gemm_scratch_bytes(int64_t M, int64_t N, int64_t K, int precision) {
  return precision == FBN_PREC_FP32 ? 0 : gemm_tc_scratch_bytes(M, N, K, precision);
}

// Times ONE GEMM configuration with CUDA events on `stream`: operands are packed once (tcgen05 precisions), then the GEMM
// kernel alone is launched `iters` times, each preceded by an (untimed) overwrite of `flush` so that operands come from HBM,
// not L2.  ms_out[0] = mean duration of the GEMM launch in milliseconds.  Synchronises the stream (benchmark helper).
extern "C" int fbn_time_gemm(const float* A, const float* Bm, float* C, int64_t M, int64_t N, int64_t K, int a_t, int b_t,
                             uint64_t kmask, int precision, void* scratch, size_t scratch_bytes, void* flush, size_t flush_bytes,
                             int iters, float* ms_out, fbn_stream_t stream) {
  FBN_REQUIRE(A && Bm && C && ms_out && iters > 0, FBN_ERR_ARG, "fbn_time_gemm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GemmArgs g;
  g.A = A; g.B = Bm; g.C = C; g.M = M; g.N = N; g.K = K; g.a_t = a_t; g.b_t = b_t; g.kmask = kmask;
  g.lda = a_t ? M : K; g.ldb = b_t ? K : N; g.ldc = N;
  uint8_t* sp = static_cast<uint8_t*>(scratch);
  if (precision != FBN_PREC_FP32) {
    const long long ra = a_t ? K : M, ca = a_t ? M : K, rb = b_t ? N : K, cb = b_t ? K : N;
    const size_t na = packed_bytes(ra, ca, precision), nb = packed_bytes(rb, cb, precision);
    FBN_REQUIRE(scratch && scratch_bytes >= na + nb, FBN_ERR_ARG, "fbn_time_gemm: scratch too small");
    RC(pack_operand(A, g.lda, ra, ca, precision, sp, ~0ull, &g.pkA, st));
    RC(pack_operand(Bm, g.ldb, rb, cb, precision, sp + na, ~0ull, &g.pkB, st));
  }
  cudaEvent_t e0, e1;
  FBN_CHECK_CUDA(cudaEventCreate(&e0));
  FBN_CHECK_CUDA(cudaEventCreate(&e1));
  RC(gemm(g, precision, nullptr, 0, st));   // warm-up (function attributes, descriptors)
  double tot = 0.0;
  for (int i = 0; i < iters; ++i) {
    if (flush) FBN_CHECK_CUDA(cudaMemsetAsync(flush, i & 0xff, flush_bytes, st));
    FBN_CHECK_CUDA(cudaEventRecord(e0, st));
    RC(gemm(g, precision, nullptr, 0, st));
    FBN_CHECK_CUDA(cudaEventRecord(e1, st));
    FBN_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FBN_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    tot += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ms_out[0] = (float)(tot / iters);
  return FBN_OK;
}

namespace fbn { void set_f16_persist_k(int k); void set_tc_pair(int on); void set_tc_persistent(int on); void set_tc_pair_persistent(int on); void set_tc_reserve_sms(int n); void set_col_chunk_mult(int m); }

// runtime knobs: "tc_pair" = 1 (default) use CTA-pair (cta_group::2) tiles for large tcgen05 GEMMs, 0 = single-CTA tiles
extern "C" int fbn_set_option(const char* name, int value) {
  FBN_REQUIRE(name != nullptr, FBN_ERR_ARG, "fbn_set_option: null name");
  if (strcmp(name, "tc_pair") == 0) { fbn::set_tc_pair(value); return FBN_OK; }
  if (strcmp(name, "f16_persist_k") == 0) { fbn::set_f16_persist_k(value); return FBN_OK; }
  if (strcmp(name, "tc_pair_persistent") == 0) { fbn::set_tc_pair_persistent(value); return FBN_OK; }
  if (strcmp(name, "ext_proj") == 0) { g_ext_proj = value; return FBN_OK; }
  if (strcmp(name, "col_chunk_mult") == 0) { fbn::set_col_chunk_mult(value); return FBN_OK; }
  if (strcmp(name, "tc_reserve_sms") == 0) { fbn::set_tc_reserve_sms(value); return FBN_OK; }
  if (strcmp(name, "side_streams") == 0) { g_use_side = value; return FBN_OK; }
  if (strcmp(name, "tc_persistent") == 0) { fbn::set_tc_persistent(value); return FBN_OK; }
  if (strcmp(name, "stage_events") == 0) { set_stage_events(value); return FBN_OK; }
  set_error("fbn_set_option: unknown option '%s'", name);
  return FBN_ERR_ARG;
}

extern "C" const char* fbn_last_error(void) { return g_err; }
extern "C" uint64_t fbn_launch_count(void) { return g_launches; }
extern "C" const char* fbn_version(void) { return "fibinet_b200 0.1 (sm_100a)"; }

extern "C" int fbn_check_device(int dev) {
  int major = 0;
  FBN_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FBN_REQUIRE(major == 10, FBN_ERR_ARCH, "device %d has compute capability %d.x; this library is sm_100a only", dev, major);
  return FBN_OK;
}

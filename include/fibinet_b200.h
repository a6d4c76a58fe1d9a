/*
 * fibinet_b200.h -- C ABI of libfibinet_b200.so
 *
 * Hand-written sm_100a CUDA implementation of the FiBiNET training / inference hot path of
 * YOUNESELBOUKNIFY/Ctr_recommendation (reference file: src/model_fibinet.py and the step body of
 * src/train_fibinet.py:113-122).  The reference has no FFI of its own (it is pure PyTorch); these
 * entry points sit *beneath* its Python module contract (build_model / forward / state_dict) and are
 * what a ctypes binding in the reference's model file would call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the library never allocates, frees or
 *     retains memory, never synchronises, and launches only on the given stream (CUDA-graph safe);
 *   - all floating point tensors are fp32, row-major, 16-byte aligned;
 *   - return value: 0 on success, negative FBN_ERR_* otherwise; fbn_last_error() gives the text;
 *   - no CPU / torch fallback exists: on a non-sm_100 device FBN_ERR_ARCH is returned.
 */
#ifndef FIBINET_B200_H_
#define FIBINET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fbn_stream_t; /* cudaStream_t */

enum {
  FBN_OK = 0,
  FBN_ERR_ARCH = -1,   /* device is not compute capability 10.x */
  FBN_ERR_SHAPE = -2,  /* unsupported sizes */
  FBN_ERR_ALIGN = -3,  /* pointer not 16-byte aligned */
  FBN_ERR_DTYPE = -4,  /* unknown index dtype */
  FBN_ERR_CUDA = -5,   /* a CUDA runtime call failed */
  FBN_ERR_ARG = -6     /* null pointer / inconsistent arguments */
};

/* dtype of the scalar index columns as the reference loader delivers them (src/dataloader.py:21-48
 * yields float64 for train/valid, int64 for test; src/model_fibinet.py:140-143 casts with .long()). */
enum { FBN_IDX_I32 = 0, FBN_IDX_I64 = 1, FBN_IDX_F64 = 2, FBN_IDX_F32 = 3 };

/* arithmetic mode of the dense contractions (bilinear + DNN tower) */
enum {
  FBN_PREC_FP32 = 0,   /* fp32 FMA (SIMT) -- exact-order reference mode */
  FBN_PREC_TF32X3 = 1, /* tcgen05 kind::tf32, 3-pass split operands, fp32 accumulate in TMEM */
  FBN_PREC_BF16 = 2,   /* tcgen05 kind::f16 bf16 operands, fp32 accumulate in TMEM */
  FBN_PREC_TF32X2 = 3, /* fbn_gemm only, K-major x K-major operands (a_t = 0, b_t = 1, K % 32 == 0): Ah*Bh as kind::tf32 plus the two
                          correction terms Al*Bh + Ah*Bl as kind::f16 bf16 MMAs -- 2 tensor-pass equivalents instead of 3,
                          ~1.4e-6 relative error; MLP-1 forward shape 517 -> 441 us.  Building block; the model path does not use it yet. */
  FBN_PREC_F16X3 = 4   /* fp32-grade result from three kind::f16 (fp16) passes: every operand tensor is scaled by ONE power of two
                          (amax -> [2^14, 2^15), found by an amax pass), split into fp16 hi | lo (the same 22 operand bits as the
                          tf32 split) and multiplied at the bf16 rate with half the operand bytes of tf32x3; the epilogue undoes
                          the scales exactly.  In the model path the long-K GEMMs (the MLP tower: forward, data- and
                          weight-gradient) run in this mode, the short-K ones (bilinear transforms, item_emb_d128 projection)
                          keep FBN_PREC_TF32X3. */
};

enum { FBN_BILINEAR_ALL = 0, FBN_BILINEAR_EACH = 1, FBN_BILINEAR_INTERACTION = 2 };

/* Fixed architecture of MM_FiBiNET (src/model_fibinet.py:92-136). */
#define FBN_D 128          /* embedding_dim == mm_input_dim */
#define FBN_F 6            /* fields [User, Like, View, ID, Image, Hist] (:112,179-182) */
#define FBN_PAIRS 15
#define FBN_K1 2688        /* (F + PAIRS) * D, MLP input (:122-123) */
#define FBN_H1 512
#define FBN_H2 256
#define FBN_SE_R 3         /* max(1, F // 2) (:13,114) */

/* Parameters of the model, in the reference's state_dict naming. */
typedef struct {
  float* item_emb;   int64_t item_rows;   /* item_emb.weight (91718,128), row 0 = padding */
  float* cate_emb;   int64_t cate_rows;   /* cate_emb.weight (11,128) */
  float* mm_w;       /* mm_proj.0.weight (128,128)  [out,in] */
  float* mm_b;       /* mm_proj.0.bias   (128) */
  float* ln_g;       /* mm_proj.1.weight (128) */
  float* ln_b;       /* mm_proj.1.bias   (128) */
  float* se_w1;      /* senet.excitation.0.weight (R,6), R = se_hidden (3) */
  float* se_b1;      /* senet.excitation.0.bias (R) */
  float* se_w2;      /* senet.excitation.2.weight (6,R) */
  float* se_b2;      /* senet.excitation.2.bias (6) */
  float* bil_w;      /* bilinear.W (128,128); for EACH: 5 matrices contiguous; INTERACTION: 15 */
  float* w1;         /* mlp.0.weight (512,2688) */
  float* b1;         /* mlp.0.bias (512) */
  float* bn1_g;      /* mlp.1.weight */
  float* bn1_b;      /* mlp.1.bias */
  float* bn1_mean;   /* mlp.1.running_mean */
  float* bn1_var;    /* mlp.1.running_var */
  float* w2;         /* mlp.4.weight (256,512) */
  float* b2;         /* mlp.4.bias */
  float* bn2_g;      /* mlp.5.weight */
  float* bn2_b;      /* mlp.5.bias */
  float* bn2_mean;   /* mlp.5.running_mean */
  float* bn2_var;    /* mlp.5.running_var */
  float* w3;         /* mlp.8.weight (1,256) */
  float* b3;         /* mlp.8.bias (1) */
  int32_t bilinear_type; /* FBN_BILINEAR_* */
  int32_t precision;     /* FBN_PREC_* */
  /* Row-sharded item table (BASELINE config 5; no counterpart in the reference, whose nn.DataParallel replicates
   * the table, src/train_fibinet.py:69-70).  n_shards == 0: item_emb is the whole (item_rows,128) table.
   * n_shards >= 1: global row g lives on rank g % n_shards at local row g / n_shards; shard[r] is the (peer-mapped,
   * see fbn_ipc_*) base of rank r's (shard_rows,128) slice and item_emb == shard[shard_rank].  The forward gather then
   * reads rows straight out of the owning GPU's HBM over NVLink -- no all-to-all, no staging copy.  In this mode the
   * workspace is sized with item_rows = 1 (fbn_workspace_bytes) and fbn_backward leaves the table gradient to
   * fbn_shard_*. */
  int32_t n_shards;
  int32_t shard_rank;
  int64_t shard_rows;
  const float* shard[16];
  /* SENetLayer hidden width max(1, 6 // reduction_ratio) (src/model_fibinet.py:13): 0 = the reference's 3 (ratio 2, hard-coded at :114);
   * 1, 2, 3, 6 = `senet_reduction` of config/fibinet_config.yaml honoured (ratios >= 4, 3, 2, 1). */
  int32_t se_hidden;
} fbn_params_t;

#define FBN_MAX_SHARDS 16

/* Gradients of the dense parameters, same naming (item_emb's gradient is handled separately). */
typedef struct {
  float* cate_emb; float* mm_w; float* mm_b; float* ln_g; float* ln_b;
  float* se_w1; float* se_b1; float* se_w2; float* se_b2; float* bil_w;
  float* w1; float* b1; float* bn1_g; float* bn1_b;
  float* w2; float* b2; float* bn2_g; float* bn2_b; float* w3; float* b3;
} fbn_grads_t;

/* One collated batch (forward(batch_dict), src/model_fibinet.py:138-146). */
typedef struct {
  int64_t batch;            /* B */
  int64_t seq_len;          /* L (<= 64); 0 or item_seq == NULL -> history field is zero (:175-176) */
  const void* item_id;      /* (B,) idx_dtype */
  const void* likes_level;  /* (B,) idx_dtype */
  const void* views_level;  /* (B,) idx_dtype */
  const void* item_seq;     /* (B,L) seq_dtype or NULL */
  const float* item_mm;     /* (B,128) fp32 item_emb_d128, or NULL when mm_table is given */
  const float* mm_table;    /* optional frozen (item_rows,128) table gathered by item_id */
  int32_t idx_dtype;        /* FBN_IDX_* of item_id / likes_level / views_level */
  int32_t seq_dtype;        /* FBN_IDX_I32 or FBN_IDX_I64 */
} fbn_batch_t;

/* Bytes of scratch the forward/backward pair needs for a batch of B rows, history length L and an
 * item table of item_rows rows.  The block must be zero-initialised once by the caller. */
size_t fbn_workspace_bytes(int64_t batch, int64_t seq_len, int64_t item_rows);

/* Byte offset of a named activation inside the workspace (for tests / debugging):
 * "X5" (B,5,128) fields 1..5 before SENET, "sgate" (B,8), "C" (B,2688) MLP input, "H1","A1" (B,512),
 * "H2","A2" (B,256), "prob","logit" (B), "ids" (B,4) int32 {item_id,likes,views,n_valid}, "seq" (B,L)
 * int32, "dV" (B,5,128), "dC" (B,2688) ...  Returns (size_t)-1 for an unknown name. */
size_t fbn_workspace_offset(int64_t batch, int64_t seq_len, int64_t item_rows, const char* name);

/* MM_FiBiNET.forward (src/model_fibinet.py:138-199).
 * train != 0: BatchNorm uses batch statistics and updates running_mean/var (momentum 0.1, unbiased
 * var), dropout p = dropout_p with either the given keep-masks (uint8 (B,512) / (B,256), test hook)
 * or an in-kernel Philox stream keyed by (seed, offset [+ *step_counter_dev << 36 when that device
 * counter is given, so a CUDA-graph replay draws fresh masks]).  prob_out: (B,) fp32 (may be NULL). */
int fbn_forward(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int train,
                float dropout_p, const uint8_t* keep_mask1, const uint8_t* keep_mask2, uint64_t seed,
                uint64_t offset, const int32_t* step_counter_dev, float* prob_out, fbn_stream_t stream);

/* The first stage of fbn_forward alone (src/model_fibinet.py:140-185): multi-field gather, history mean
 * pooling, item_emb_d128 projection + LayerNorm + ReLU, field stack and SENET; writes the re-weighted
 * fields into the "C" workspace tensor (columns 128..767) and, if save != 0, the tensors backward needs. */
int fbn_embed_forward(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int save,
                      fbn_stream_t stream);

/* Backward of the forward above given dL/dprob (B,) (what autograd hands over after
 * torch.nn.BCELoss, src/train_fibinet.py:115-116).  Dense-parameter gradients are WRITTEN to g.
 * The item_emb gradient is produced as a dense (item_rows,128) table in item_grad:
 *   zero_fill != 0 : every row written (untouched rows = 0) -- the nn.Embedding(sparse=False) contract;
 *   zero_fill == 0 : only touched rows written; row_touched (item_rows,) int32 (optional) receives the
 *                    occurrence count per row (consumed by fbn_adam_table).
 * grad_sumsq (2,) receives sum(g^2) of [dense grads (if dense_grad_flat != NULL: the n floats at that
 * address, normally the flat buffer all pointers of g point into), item_emb grad] for the global clip. */
int fbn_backward(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int train,
                 float dropout_p, const float* dprob, const fbn_grads_t* g, const float* dense_grad_flat,
                 int64_t dense_grad_n, float* item_grad, int32_t* row_touched, int zero_fill, int index_ready,
                 float* grad_sumsq, fbn_stream_t stream);

/* fbn_backward in phases, for a data-parallel host that overlaps its gradient collectives with the weight-gradient GEMMs
 * (replaces the grad reduce of nn.DataParallel, src/train_fibinet.py:69-70; BASELINE north_star (5)):
 *   FBN_BWD_CHAIN : the data-gradient chain down to the embedding-table rows (item_grad / row_touched / grad_sumsq[1] as above,
 *                   plus the gradients computed along the chain: mlp.8.*, mlp.5.*, mlp.1.*);
 *   FBN_BWD_LEAF1 : mlp.0.weight / mlp.0.bias;   FBN_BWD_LEAF2 : every other parameter gradient.
 * A leaf requested together with CHAIN runs beside it on the library's side stream (as in fbn_backward); a leaf requested by a later
 * call on the same workspace runs on `stream` (dprob is only read by CHAIN).  phases == CHAIN|LEAF1|LEAF2 is fbn_backward without
 * the dense-gradient norm. */
enum { FBN_BWD_CHAIN = 1, FBN_BWD_LEAF1 = 2, FBN_BWD_LEAF2 = 4 };
int fbn_backward_phase(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int train,
                       float dropout_p, const float* dprob, const fbn_grads_t* g, float* item_grad, int32_t* row_touched,
                       int zero_fill, int index_ready, float* grad_sumsq, int phases, fbn_stream_t stream);

/* The occurrence index fbn_backward needs for the embedding gradient (stable sort of the B + B*L row ids, per-row
 * counts and offsets).  It depends on the batch ids only: run it on a second stream concurrently with fbn_forward and
 * pass index_ready = 1 to fbn_backward (same row_touched pointer), or pass index_ready = 0 and let fbn_backward build it. */
int fbn_embed_index(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, int32_t* row_touched,
                    fbn_stream_t stream);

/* Fused BCELoss(mean) forward + loss_scale * d(loss)/d(prob) (torch semantics: log clamped at -100,
 * backward divides by max(p(1-p),1e-12)); loss_out (1,) / dprob_out (B,) may be NULL. */
int fbn_bce_loss(const float* prob, const float* labels, int64_t batch, float loss_scale, float* loss_out,
                 float* dprob_out, fbn_stream_t stream);
/* The same with a multi-block kernel (fixed-order final sum: deterministic) for large batches; scratch: fbn_bce_scratch_bytes()
 * bytes, 8-byte aligned, ZERO-INITIALISED once by the caller (it holds the arrival counter, which every launch re-arms). */
size_t fbn_bce_scratch_bytes(void);
int fbn_bce_loss_ws(const float* prob, const float* labels, int64_t batch, float loss_scale, float* loss_out,
                    float* dprob_out, void* scratch, size_t scratch_bytes, fbn_stream_t stream);

/* clip_grad_norm_ coefficient (src/train_fibinet.py:119): coef = min(1, max_norm/(sqrt(sum)+1e-6)),
 * sum = sumsq[0] + ... + sumsq[n-1].  out (2,) = {total_norm, coef}. */
int fbn_clip_coef(const float* sumsq, int n, float max_norm, float* out, fbn_stream_t stream);

/* Hyper-parameters of torch.optim.Adam as the reference builds it (src/train_fibinet.py:78):
 * L2 weight decay folded into the gradient, bias corrections from the current beta1 (OneCycleLR
 * cycles it, :84-92).  step is the 1-based step count.  Every Adam entry point takes either this
 * struct (host values) or hyper_dev, a device array of 12 floats {lr, beta1, beta2, eps, wd, lr/(1-beta1^t),
 * sqrt(1-beta2^t), t, 1-beta1, 1-beta2, decay multiplier, 0} (CUDA-graph replay; filled by fbn_onecycle_hyper or by the host).
 * one_minus_beta1/2: torch evaluates `1 - beta` in double precision before it becomes an fp32 kernel scalar (for
 * beta2 = 0.999 that differs from 1.0f - 0.999f by 1.3e-5 relative); pass those doubles rounded to float, or 0 to
 * let the library derive them from the fp32 betas.
 * decoupled != 0: torch.optim.AdamW semantics instead (p *= 1 - lr*wd, then Adam on the raw gradient) -- what `optimizer: adamw`
 * in config/fibinet_config.yaml names but the reference never builds (src/train_fibinet.py:78 constructs torch.optim.Adam);
 * hyper_dev[10] = 1 - lr*wd (0 = coupled L2). */
typedef struct { float lr, beta1, beta2, eps, weight_decay; int32_t step; float one_minus_beta1, one_minus_beta2; int32_t decoupled; } fbn_adam_t;

/* Dense-exact Adam over the whole embedding table, consuming the gradient rows produced by
 * fbn_backward: g = (touched ? grad[row] : 0) * coef + wd*p, then the Adam update, for EVERY row
 * (SURVEY fact 6).  coef is read from device memory clip[1] (clip == NULL -> 1).
 * row_touched == NULL -> grad is read for every row. */
int fbn_adam_table(float* p, float* m, float* v, const float* grad, const int32_t* row_touched,
                   int64_t rows, const float* clip, const fbn_adam_t* h, const float* hyper_dev,
                   fbn_stream_t stream);

/* Flat dense Adam over n contiguous fp32 elements (all dense parameters live in one buffer; n % 4 == 0). */
int fbn_adam_dense(float* p, float* m, float* v, const float* grad, int64_t n, const float* clip,
                   const fbn_adam_t* h, const float* hyper_dev, fbn_stream_t stream);

/* Adagrad row update (BASELINE north_star (2): the sorted-segment sums of fbn_backward "fused into the Adam/Adagrad row update").
 * Replaces the optimizer site src/train_fibinet.py:78 when the user asks for Adagrad; the reference itself builds Adam, so this is
 * an extension whose oracle is torch.optim.Adagrad (single-tensor path of torch/optim/adagrad.py):
 *   g = grad * coef + wd * p ; clr = lr / (1 + (step - 1) * lr_decay) ; state_sum += g * g ; p -= clr * g / (sqrt(state_sum) + eps).
 * step is the 1-based step count.  hyper_dev (CUDA-graph replay): the same 12-float device array as Adam with
 * [0] = clr, [3] = eps, [4] = wd.  With wd == 0 an untouched row is the identity and is neither read nor written. */
typedef struct { float lr, lr_decay, eps, weight_decay; int32_t step; } fbn_adagrad_t;
int fbn_adagrad_table(float* p, float* state_sum, const float* grad, const int32_t* row_touched, int64_t rows,
                      const float* clip, const fbn_adagrad_t* h, const float* hyper_dev, fbn_stream_t stream);
int fbn_adagrad_dense(float* p, float* state_sum, const float* grad, int64_t n, const float* clip,
                      const fbn_adagrad_t* h, const float* hyper_dev, fbn_stream_t stream);

/* OneCycleLR(cos, two phases, cycle_momentum=True) as built at src/train_fibinet.py:84-92, evaluated on
 * the device: reads *step_counter (optimizer steps taken so far), writes hyper_dev[0..7] for the next
 * step and increments the counter. */
int fbn_onecycle_hyper(int32_t* step_counter, int total_steps, float max_lr, float pct_start,
                       float div_factor, float final_div_factor, float base_momentum, float max_momentum,
                       float beta2, float eps, float weight_decay, float* hyper_dev, fbn_stream_t stream);

/* ---- row-sharded item table: gradient exchange over peer memory --------------------------------------------------
 * One process per GPU.  Every rank owns an "exchange block" of fbn_shard_xchg_bytes(cap) bytes (cap = the largest
 * B*(1+L) of any rank) that all peers map with fbn_ipc_open.  Per step:
 *   fbn_shard_index      ids only: occurrence keys (owner, local row), stable radix sort, run heads -> the rank's unique
 *                        rows grouped by owner (may run concurrently with fbn_forward);
 *   fbn_backward         with item_grad == NULL (dX rows stay in the workspace: "dXitem", "dXhist");
 *   fbn_shard_local_sum  one warp per unique row sums its occurrences in source order into the exchange block;
 *   -- any inter-rank barrier (the dense-gradient all-reduce is one) --
 *   fbn_shard_merge      the OWNER pulls, over NVLink, every peer's partial rows for its slice, merges the N sorted lists
 *                        by rank (stable: contributions are added in rank order, bitwise reproducible) and writes either a
 *                        dense (shard_rows,128) gradient + touched flags (dense-exact Adam via fbn_adam_table) or a compact
 *                        (row, gradient) list (lazy row Adam via fbn_shard_adam_rows); sumsq_out (1,) = sum(g^2) of the slice;
 *   -- all-reduce of the sum of squares (clip) = barrier: the exchange blocks may be rewritten --
 *   fbn_adam_table / fbn_shard_adam_rows on the local slice;
 *   -- barrier before the next forward reads the updated rows remotely --                                              */
typedef struct {
  int32_t n_shards, rank;
  int64_t item_rows;    /* global rows V */
  int64_t shard_rows;   /* ceil(V / n_shards) */
  int64_t cap;          /* occurrence capacity of every rank's exchange block */
  int64_t merge_cap;    /* capacity (items) of the owner-side merge lists; <= n_shards * cap */
  void* xchg[FBN_MAX_SHARDS]; /* peer-mapped exchange blocks; xchg[rank] is the local one */
} fbn_shard_plan_t;
size_t fbn_shard_xchg_bytes(int64_t cap);
size_t fbn_shard_ws_bytes(int64_t cap, int64_t merge_cap, int n_shards, int64_t shard_rows);
int fbn_shard_index(const fbn_shard_plan_t* s, const fbn_batch_t* b, void* sws, size_t sws_bytes, fbn_stream_t stream);
int fbn_shard_local_sum(const fbn_shard_plan_t* s, const fbn_batch_t* b, const float* dXitem, const float* dXhist, void* sws,
                        size_t sws_bytes, fbn_stream_t stream);
/* dense_grad (shard_rows,128) + touched (shard_rows,) int32 [zeroed inside], or both NULL for the compact list kept in sws */
int fbn_shard_merge(const fbn_shard_plan_t* s, void* sws, size_t sws_bytes, float* dense_grad, int32_t* touched,
                    float* sumsq_out, fbn_stream_t stream);
/* Lazy row Adam over the compact list of the last fbn_shard_merge: only touched rows move (g = grad*coef + wd*p, torch
 * single-tensor op order, bias corrections from the global step) -- extension, the reference's Adam is dense. */
int fbn_shard_adam_rows(const fbn_shard_plan_t* s, void* sws, size_t sws_bytes, float* p, float* m, float* v, const float* clip,
                        const fbn_adam_t* h, const float* hyper_dev, fbn_stream_t stream);
/* debug / tests: copies {U, owner_start[0..n], T, Um, overflow flag} of the last step to host_out (>= 24 int32); synchronises */
int fbn_shard_stats(const fbn_shard_plan_t* s, void* sws, size_t sws_bytes, int32_t* host_out, fbn_stream_t stream);

/* CUDA IPC plumbing for the peer-mapped buffers above (legacy cudaIpc*: memory must come from cudaMalloc, e.g. torch's
 * default caching allocator without expandable_segments).  export: handle (64 bytes) of the allocation containing dev_ptr
 * and dev_ptr's byte offset inside it.  open: maps the allocation into this process (peer access enabled lazily) and
 * returns its base; open each distinct handle once per process.  close: unmaps a base returned by open. */
int fbn_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out);
int fbn_ipc_open(const void* handle, void** base_out);
int fbn_ipc_close(void* base);

/* ---- F-field FiBiNET (BASELINE config 5: "scaled synthetic FiBiNET: 40 fields"): the building blocks of general.py --------------
 * The reference hard-wires six fields (src/model_fibinet.py:112,179-182); its blocks are generic in F and are what is exposed
 * here: F embedding lookups or mean-pooled id bags (tables stored back to back in ONE (sum of vocabularies, 128) matrix), SENetLayer /
 * BilinearInteraction for F <= 64 (fbn_senet_*, fbn_bilinear_*_ld) and the MLP tower for any input width K1 = (F + F(F-1)/2) * 128.
 * Oracle: oracle/fibinet_general.py, pinned against the reference's own SENetLayer / BilinearInteraction classes. */

/* Field lookups (nn.Embedding, src/model_fibinet.py:155-159) and masked mean-pooled bags (the reference's item_seq pooling, :165-174).
 * desc: (F,5) int64 on the DEVICE, per field {first row of its table inside `table`, vocabulary, first id column, bag length, padding id
 * (-1 = none)}; several fields may share a table.  ids: (B, id_cols) int32 / int64, one column per id of every field (id_cols = sum of
 * the bag lengths).  x (B,F,128): the looked-up row (bag length 1) or the mean of the rows whose id is not the padding id; count (B,F):
 * max(number of such ids, 1).  flag (4) int32 on the device: flag[0] is set when an id is outside [0, vocab) (the kernel clamps; the
 * host raises IndexError). */
int fbn_fields_gather(const float* table, const int64_t* desc, const void* ids, int idx_dtype, int64_t batch, int fields, int id_cols,
                      float* x, float* count, int32_t* flag, fbn_stream_t stream);
/* Dense gradient of the lookups above (embedding_dense_backward): grad (rows,128) = scatter-add of dx[b][f] / count[b][f] to every
 * non-padding id of field f, by the same deterministic sorted-segment sum as the item table (stable sort of the B * id_cols global
 * rows, one warp per row in source order, hot rows chunked).  dx is divided by count IN PLACE.  col_field: (id_cols) int32 on the
 * device, the field of every id column.  zero_fill / row_touched / sumsq_out (1,) as in fbn_backward.
 * scratch: fbn_fields_scatter_bytes(batch, id_cols, rows). */
size_t fbn_fields_scatter_bytes(int64_t batch, int id_cols, int64_t rows);
int fbn_fields_scatter(float* dx, const float* count, const int64_t* desc, const int32_t* col_field, const void* ids, int idx_dtype,
                       int64_t batch, int fields, int id_cols, int64_t rows, float* grad, int32_t* row_touched, int zero_fill,
                       float* sumsq_out, void* scratch, size_t scratch_bytes, fbn_stream_t stream);

/* The MLP tower of the reference for an arbitrary input width k1 (src/model_fibinet.py:125-136,197-199): Linear(k1,512) -> BatchNorm1d ->
 * ReLU -> Dropout -> Linear(512,256) -> BatchNorm1d -> ReLU -> Dropout -> Linear(256,1) -> sigmoid, and its backward.  Only the MLP
 * fields of params / grads are read (w1 is (512,k1)).  c: (B,k1) fp32 MLP input; dc: (B,k1) gradient w.r.t. it.  The workspace
 * (fbn_tower_workspace_bytes, zero-initialised once) keeps the activations between the two calls.  Semantics of train / dropout /
 * masks / seed / step counter / dprob as in fbn_forward / fbn_backward; the same tcgen05 / SIMT GEMM back ends (params->precision). */
size_t fbn_tower_workspace_bytes(int64_t batch, int64_t k1);
/* byte offset of "H1" / "A1" (B,512), "H2" / "A2" (B,256), "logit" / "prob" / "dlogit" (B), "dH1", "dH2" inside it (tests); (size_t)-1 = unknown */
size_t fbn_tower_workspace_offset(int64_t batch, int64_t k1, const char* name);
int fbn_tower_forward(const fbn_params_t* p, const float* c, int64_t batch, int64_t k1, void* ws, size_t ws_bytes, int train,
                      float dropout_p, const uint8_t* keep_mask1, const uint8_t* keep_mask2, uint64_t seed, uint64_t offset,
                      const int32_t* step_counter_dev, float* prob_out, float* logit_out, fbn_stream_t stream);
int fbn_tower_backward(const fbn_params_t* p, const float* c, int64_t batch, int64_t k1, void* ws, size_t ws_bytes, int train,
                       float dropout_p, const float* dprob, const fbn_grads_t* g, float* dc, fbn_stream_t stream);

/* sum of squares of n floats, deterministic; partial needs fbn_sumsq_partial_floats(n) floats;
 * out (1,) overwritten. */
int fbn_sumsq(const float* x, int64_t n, float* partial, float* out, fbn_stream_t stream);
size_t fbn_sumsq_partial_floats(int64_t n);

/* Stand-alone SENetLayer (src/model_fibinet.py:5-35): x (B,F,D) -> y (B,F,D), gate (B,F).
 * F <= 64, hidden <= 64, D % 4 == 0. */
int fbn_senet_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                  int64_t batch, int fields, int dim, int hidden, float* y, float* gate, fbn_stream_t stream);
/* scratch: fbn_senet_scratch_bytes(batch, fields, hidden) bytes */
int fbn_senet_bwd(const float* x, const float* gate, const float* w1, const float* b1, const float* w2,
                  const float* dy, int64_t batch, int fields, int dim, int hidden, float* dx, float* dw1,
                  float* db1, float* dw2, float* db2, void* scratch, size_t scratch_bytes, fbn_stream_t stream);
size_t fbn_senet_scratch_bytes(int64_t batch, int fields, int hidden);

/* Stand-alone BilinearInteraction (src/model_fibinet.py:37-89): v (B,F,D) -> p (B,F(F-1)/2,D).
 * w: ALL (D,D); EACH (F-1,D,D); INTERACTION (F(F-1)/2,D,D).  D % 4 == 0. */
int fbn_bilinear_fwd(const float* v, const float* w, int type, int64_t batch, int fields, int dim,
                     float* p, void* scratch, size_t scratch_bytes, int precision, fbn_stream_t stream);
int fbn_bilinear_bwd(const float* v, const float* w, const float* dp, int type, int64_t batch, int fields,
                     int dim, float* dv, float* dw, void* scratch, size_t scratch_bytes, int precision,
                     fbn_stream_t stream);

/* Strided variants for the F-field model (general.py): the pair products are written straight into / read straight from the
 * (B, (F+P)*dim) MLP-input buffer (row strides ldp / lddp in floats), and dv starts from dv_init (row stride ld_init; NULL = 0) --
 * the gradient that reached the fields directly through the concat (src/model_fibinet.py:191-194). */
int fbn_bilinear_fwd_ld(const float* v, const float* w, int type, int64_t batch, int fields, int dim, float* p, int64_t ldp,
                        void* scratch, size_t scratch_bytes, int precision, fbn_stream_t stream);
int fbn_bilinear_bwd_ld(const float* v, const float* w, const float* dp, int64_t lddp, const float* dv_init, int64_t ld_init,
                        int type, int64_t batch, int fields, int dim, float* dv, float* dw, void* scratch, size_t scratch_bytes,
                        int precision, fbn_stream_t stream);
size_t fbn_bilinear_scratch_bytes(int64_t batch, int fields, int dim, int type);

/* C[M,N] = op(A) * op(B) (+ bias[N]) in the given precision -- exposed for tests and benches.
 * a_t == 0: A is (M,K) row-major lda; a_t != 0: A is stored (K,M) row-major lda.
 * b_t == 0: B is (K,N) row-major ldb; b_t != 0: B is stored (N,K) row-major ldb (nn.Linear weight). */
int fbn_gemm(const float* A, const float* B, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
             int64_t lda, int64_t ldb, int64_t ldc, int a_t, int b_t, int precision, void* scratch,
             size_t scratch_bytes, fbn_stream_t stream);
/* scratch the tcgen05 precisions need for their packed operands (0 for FBN_PREC_FP32) */
size_t fbn_gemm_scratch_bytes(int64_t M, int64_t N, int64_t K, int precision);

/* Benchmark helper: mean CUDA-event duration (ms, host pointer ms_out) of the GEMM kernel alone for one configuration --
 * operands packed once, `iters` timed launches each preceded by an untimed overwrite of `flush` (L2 eviction).
 * A is (M,K) [a_t=0] or (K,M) [a_t=1] contiguous; B is (K,N) [b_t=0] or (N,K) [b_t=1] contiguous; kmask as in the MLP path. */
int fbn_time_gemm(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int a_t, int b_t, uint64_t kmask,
                  int precision, void* scratch, size_t scratch_bytes, void* flush, size_t flush_bytes, int iters, float* ms_out,
                  fbn_stream_t stream);
/* Benchmark helper: mean CUDA-event duration (ms) of one stage of the forward pass exactly as fbn_forward launches it, after a
 * complete fbn_forward on the same workspace: "embed", "bil_gemm", "bil_pairs" (f16x3: the amax + split passes that write the MLP input), "mlp1", "mlp1_dgrad", "mlp1_wgrad"
 * (the last two after an fbn_backward).  `flush` is overwritten before
 * every timed launch (L2 eviction).  Synchronises the stream. */
int fbn_time_stage(const fbn_params_t* p, const fbn_batch_t* b, void* ws, size_t ws_bytes, const char* stage, void* flush,
                   size_t flush_bytes, int iters, float* ms_out, fbn_stream_t stream);
/* With fbn_set_option("stage_events", 1) (eager launches only, not under graph capture) fbn_forward / fbn_backward record a CUDA
 * event at every stage boundary; this writes "stage<TAB>milliseconds" lines for the calls since the last report.  Synchronises. */
int fbn_stage_report(char* buf, size_t buf_bytes);
/* runtime knobs: "tc_pair" (1 = CTA-pair 256x256 tcgen05 tiles for large GEMMs [default], 0 = single-CTA 128x128),
 * "tc_persistent" (0 = per-launch heuristic [default]: persistent tile loop for short-K / bf16 epilogue-bound GEMMs,
 * 1 = always, -1 = never), "tc_pair_persistent" (1 = the CTA pairs walk a tile list [default], 0 = one tile per cluster),
 * "f16_persist_k" (K threshold below which an f16x3 GEMM takes the single-CTA persistent loop; 0 = the tf32x3 rule [default]),
 * "tc_reserve_sms" (SMs the persistent GEMMs leave to a concurrent collective), "side_streams", "ext_proj", "col_chunk_mult",
 * "stage_events" (see fbn_stage_report) */
int fbn_set_option(const char* name, int value);
/* number of kernels this library has launched so far in this process (host-side counter) */
uint64_t fbn_launch_count(void);
const char* fbn_last_error(void);
const char* fbn_version(void);
/* 0 if device `dev` can run the library (compute capability 10.x), FBN_ERR_ARCH otherwise. */
int fbn_check_device(int dev);

#ifdef __cplusplus
}
#endif
#endif /* FIBINET_B200_H_ */

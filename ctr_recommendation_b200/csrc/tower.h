// Internal launchers (see tower.cu / embed.cu / embbwd.cu / optim.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fbn {

// Optional second destination of a producer kernel: the same values in tcgen05 operand format (tf32 hi|lo or bf16),
// element (r, c) at r * pitch + c.  mode 0 = disabled.  Lets a tensor be converted by the kernel that produces it instead
// of a separate pack pass.
struct PackDst {
  void* base = nullptr;
  long long pitch = 0, lo_off = 0;
  int mode = 0;
  // f16x3: the tensor is split AFTER it is complete (its scale needs the maximum of |x|); the producer only publishes its
  // per-block maxima into the record that follows the packed region (common.cuh), launched with at most F16_AMAX_BLOCKS blocks
  float* tail = nullptr;
};

struct DropArgs {
  float p = 0.f;                 // 0 -> no dropout
  const uint8_t* mask = nullptr; // optional explicit keep-mask (B,N) uint8
  uint64_t seed = 0, offset = 0, stream = 0;
  const int32_t* step_dev = nullptr;  // optional device step counter mixed into the offset (CUDA-graph replay)
};

int col_chunks(long long B, int N);
int colsum(const float* X, long long B, int N, float* partial, float* out, cudaStream_t st);
int colprod2(const float* X, const float* Y, long long B, int N, float* partial, float* out_xy, float* out_x, cudaStream_t st);
int bn_train_stats(const float* H, long long B, int N, float* partial, float* mean, float* rstd, float* run_mean, float* run_var,
                   cudaStream_t st);
int bn_eval_stats(const float* run_mean, const float* run_var, int N, float* mean, float* rstd, cudaStream_t st);
int bn_act(const float* H, const float* mean, const float* rstd, const float* g, const float* b, long long B, int N,
           const DropArgs& d, float* A, PackDst pk, cudaStream_t st);
int head_fwd(const float* H, const float* mean, const float* rstd, const float* g, const float* b, const float* w3, const float* b3,
             long long B, const DropArgs& d, float* A, float* logit, float* prob, cudaStream_t st);
int head_bwd_stats(const float* dprob, const float* prob, const float* A2, const float* Hd2, const float* mean, const float* rstd,
                   const float* w3, long long B, float scale, float* partial, float* dlogit, float* dgamma, float* dbeta, float* dw3,
                   float* db3, cudaStream_t st);
int bn_bwd_stats(const float* dA, const float* A, const float* Hd, const float* mean, const float* rstd, long long B, int N,
                 float scale, float* partial, float* dgamma, float* dbeta, cudaStream_t st);
int bn_bwd_apply(const float* dA, const float* dlogit, const float* w3, const float* A, const float* Hd, const float* mean,
                 const float* rstd, const float* g, const float* dgamma, const float* dbeta, long long B, int N, float scale,
                 int train, float* dH, PackDst pk, cudaStream_t st);
int bilinear_pairs_fwd(int type, float* C, const float* T, long long B, PackDst pk, cudaStream_t st);
// f16x3: the MLP input as fp16 hi|lo under one scale, written in one go from the field blocks of C and the transforms T -- an amax
// pass (no writes) and a split pass that recomputes the pair products, so the fp32 pair blocks never exist (dst16: hi part,
// lo part lo_off halves later, pitch K1; tail: the region's record)
int bilinear_pairs_mlp16(int type, const float* C, const float* T, long long B, void* dst16, long long lo_off, float* tail, cudaStream_t st);
int bilinear_pairs_bwd(int type, const float* C, const float* T, const float* dC, long long B, float* dT, float* dV, PackDst pk,
                       cudaStream_t st);
int reduce_splits(const float* partial, int parts, long long M, long long N, long long part_stride, unsigned long long nmask, float* out,
                  cudaStream_t st);

// embed.cu
int launch_reduce_partials(const float* partial, float* out, int parts, long long n, int accumulate, cudaStream_t st);
int launch_senet_param_grads(const float* sestat, long long B, int R, float* partial, float* dw1, float* db1, float* dw2, float* db2,
                             cudaStream_t st);
int embed_bwd_blocks(long long B);

// embbwd.cu : deterministic sorted-segment embedding backward
struct EmbGradArgs {
  const void* item_id; int idx_dtype;   // raw batch columns (the index does not depend on the forward pass)
  const void* seq; int seq_dtype;       // (B,L) or nullptr
  long long B; int L; long long rows;
  const float* dXitem;     // (B,128)
  const float* dXhist;     // (B,128), already divided by the history count
  int32_t* keys_in; int32_t* keys_out; int32_t* vals_in; int32_t* vals_out;
  int32_t* row_count;      // (rows) occurrences per table row (output)
  int32_t* row_off;        // (rows+1) exclusive scan (output)
  void* cub_tmp; size_t cub_bytes;
  float* grad;             // (rows,128) table gradient rows
  int zero_fill;           // write zeros to untouched rows (dense .grad contract)
  float* sumsq_partial;    // per-CTA partial sums of squares
  float* sumsq_out;        // (1)
};
int emb_index(const EmbGradArgs& a, cudaStream_t st);
int emb_rows(const EmbGradArgs& a, cudaStream_t st);
size_t emb_sort_temp_bytes(long long n, long long rows);
int emb_grad_partial_count(long long rows);

// optim.cu
int sumsq(const float* x, long long n, float* partial, float* out, cudaStream_t st);
int sumsq_partial_count(long long n);
float one_minus(float beta);   // 1 - beta the way torch evaluates it (double arithmetic on the decimal literal)

}  // namespace fbn

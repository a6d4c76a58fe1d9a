// Argument blocks of the fused embedding kernels (embed.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tower.h"

namespace fbn {

struct EmbedFwdArgs {
  const float* item_emb; const float* cate_emb;
  const float* mm_w; const float* mm_b; const float* ln_g; const float* ln_b;
  const float* se_w1; const float* se_b1; const float* se_w2; const float* se_b2;
  const void* item_id; const void* likes; const void* views; const void* seq;
  const float* item_mm; const float* mm_table;
  const float* yproj;  // optional (B,128): item_emb_d128 x mm_w^T + mm_b computed beforehand (tcgen05 GEMM); the kernel then skips its own projection
  int idx_dtype, seq_dtype;
  long long B; int L; long long item_rows; int cate_rows;
  int save;            // write the tensors backward needs
  int32_t* ids; int32_t* seq32;
  int32_t* idflag;     // [0] item_id, [1] likes/views, [2] item_seq out of range (sticky; torch raises IndexError for these)
  float* X5; float* sgate; float* xhat; float* xmm; float* rstd; float* cnt; float* C;
  PackDst pkC;         // packed copy of the field blocks of C
  PackDst pkX;         // packed copy of the item_emb_d128 rows (B,128)
  int se_r;            // SENET hidden width: max(1, 6 // reduction_ratio) in {1, 2, 3, 6}
  int nshard;          // > 0: row-sharded item table, row g = shard[g % nshard] + (g / nshard) * 128 (peer-mapped pointers)
  const float* shard[FBN_MAX_SHARDS];
};

struct EmbedBwdArgs {
  const float* dV;      // (B,5,128) gradient w.r.t. SENET output fields 1..5
  const float* X5; const float* sgate; const float* xhat; const float* rstd; const float* cnt;
  const int32_t* ids;
  const float* se_w1; const float* se_b1; const float* se_w2; const float* ln_g;
  long long B; int cate_rows; int se_r;
  float* dXitem; float* dXhist; float* dln; float* dy; float* sestat;
  PackDst pkdy;
  float* cate_partial;  // (gridDim.x, cate_rows, 128)
};

int launch_embed_senet_fwd(const EmbedFwdArgs& a, cudaStream_t st);
int launch_embed_senet_bwd(const EmbedBwdArgs& a, int blocks, cudaStream_t st);

}  // namespace fbn

// G1 + S1: fused multi-field embedding gather, history mean-pooling, item_emb_d128 projection
// (Linear 128->128 + LayerNorm + ReLU), field stack and SENET squeeze-excitation, and its backward.
//
// Reference: src/model_fibinet.py:140-185 (forward), autograd of the same (backward).
// One warp owns one sample row at a time: D = 128 fp32 = 32 lanes x float4, so every table row,
// every field of a sample and every output row is a single fully coalesced 512-byte warp access.
// The (B,20,128) history tensor the reference materialises (42 MB at B=4096) never exists here:
// rows are summed in registers in sequence order.
#include <algorithm>

#include "common.cuh"
#include "embed_args.h"
#include "tower.h"

namespace fbn {

constexpr int EMB_WARPS = 8;
constexpr int EMB_SPW = 2;  // samples per warp per iteration (register-blocks the projection; 4 measured: 250 vs 254 us at B = 65536 but 42 vs 31 us at 4096 -- half the CTAs)


// smem: Wt[128][128] (k-major copy of mm_w so lane j reads W[4j..4j+3][k] as one float4)
//       xs[EMB_WARPS][128][EMB_SPW]
// row g of the item table: the local (replicated) table, or -- row-sharded mode -- the owner's slice, read over NVLink
template <bool SHARDED>
__device__ __forceinline__ const float* item_row(const EmbedFwdArgs& a, long long id) {
  if (!SHARDED) return a.item_emb + id * D;
  const unsigned g = (unsigned)id, n = (unsigned)a.nshard;
  return a.shard[g % n] + (long long)(g / n) * D;
}

// EXTP: the item_emb_d128 projection y = x W^T + b was computed for the whole batch by a tcgen05 GEMM beforehand (a.yproj, (B,128));
// the kernel then needs neither the 64 KB copy of W in shared memory nor the SIMT projection loop, and (EXTP ? 3 : 2) CTAs fit an SM.
template <bool SHARDED, int SE_R, bool EXTP>
__global__ void __launch_bounds__(EMB_WARPS * 32, EXTP ? 3 : 2) embed_senet_fwd_kernel(EmbedFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Wt = smem;
  float* xs_all = smem + D * D;
  __shared__ float s_se[SE_R * NF + SE_R + NF * SE_R + NF];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // k-major copy of mm_w, float4 columns XOR-swizzled by k so that the transposing stores below are 4-way instead of 32-way
  // bank-conflicted (the prologue is paid by every CTA and dominates small batches)
  if (!EXTP) {
#pragma unroll 8
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
      const int j = i >> 7, k = i & 127;  // mm_w[j][k], coalesced read
      Wt[k * D + ((((j >> 2) ^ (k & 31)) << 2) | (j & 3))] = __ldg(a.mm_w + i);
    }
  }
  if (threadIdx.x < SE_R * NF) s_se[threadIdx.x] = a.se_w1[threadIdx.x];
  if (threadIdx.x < SE_R) s_se[SE_R * NF + threadIdx.x] = a.se_b1[threadIdx.x];
  if (threadIdx.x < NF * SE_R) s_se[SE_R * NF + SE_R + threadIdx.x] = a.se_w2[threadIdx.x];
  if (threadIdx.x < NF) s_se[SE_R * NF + SE_R + NF * SE_R + threadIdx.x] = a.se_b2[threadIdx.x];
  __syncthreads();
  const float* w1 = s_se; const float* b1 = s_se + SE_R * NF;
  const float* w2 = b1 + SE_R; const float* b2 = w2 + NF * SE_R;
  float* xs = EXTP ? nullptr : xs_all + warp * (D * EMB_SPW);
  const float4 bias = EXTP ? f4(0.f) : ld4(a.mm_b + 4 * lane), gam = ld4(a.ln_g + 4 * lane), bet = ld4(a.ln_b + 4 * lane);

  const long long ngroups = (a.B + EMB_SPW - 1) / EMB_SPW;
  for (long long g = (long long)blockIdx.x * EMB_WARPS + warp; g < ngroups; g += (long long)gridDim.x * EMB_WARPS) {
    float4 f_like[EMB_SPW], f_view[EMB_SPW], f_item[EMB_SPW], f_hist[EMB_SPW];
    float cntv[EMB_SPW];
#pragma unroll
    for (int s = 0; s < EMB_SPW; ++s) {
      const long long b = g * EMB_SPW + s;
      f_like[s] = f_view[s] = f_item[s] = f_hist[s] = f4(0.f);
      cntv[s] = 1.f;
      float4 xm = f4(0.f);
      if (b < a.B) {
        long long iid = 0, lk = 0, vw = 0;
        if (lane == 0) iid = load_index(a.item_id, a.idx_dtype, b);
        if (lane == 1) lk = load_index(a.likes, a.idx_dtype, b);
        if (lane == 2) vw = load_index(a.views, a.idx_dtype, b);
        iid = __shfl_sync(0xffffffffu, iid, 0);
        lk = __shfl_sync(0xffffffffu, lk, 1);
        vw = __shfl_sync(0xffffffffu, vw, 2);
        // out-of-range ids are an IndexError in torch (nn.Embedding, ref :155-159): flag them for the host, which raises, and
        // clamp so the kernel itself cannot fault
        if (lane == 0 && (iid < 0 || iid >= a.item_rows)) a.idflag[0] = 1;
        if (lane == 1 && (lk < 0 || lk >= a.cate_rows || vw < 0 || vw >= a.cate_rows)) a.idflag[1] = 1;
        iid = min(max(iid, 0LL), a.item_rows - 1);
        lk = min(max(lk, 0LL), (long long)a.cate_rows - 1);
        vw = min(max(vw, 0LL), (long long)a.cate_rows - 1);
        f_like[s] = ld4(a.cate_emb + lk * D + 4 * lane);
        f_view[s] = ld4(a.cate_emb + vw * D + 4 * lane);
        f_item[s] = ld4(item_row<SHARDED>(a, iid) + 4 * lane);
        if (!EXTP) {
          if (a.item_mm) {
            xm = ld4s(a.item_mm + b * D + 4 * lane);
          } else {
            xm = ld4(a.mm_table + iid * D + 4 * lane);
            if (a.save) st4(a.xmm + b * D + 4 * lane, xm);   // wgrad of mm_proj.0.weight needs the gathered rows
          }
          if (a.save && a.pkX.mode) store_packed4(a.pkX.base, a.pkX.lo_off, a.pkX.mode, b * D + 4 * lane, xm);
        }
        int nvalid = 0;
        float4 acc = f4(0.f);
        if (a.seq != nullptr) {
          for (int l0 = 0; l0 < a.L; l0 += 32) {
            int myid = 0;
            if (l0 + lane < a.L) {
              long long v = load_index(a.seq, a.seq_dtype, b * a.L + l0 + lane);
              if (v < 0 || v >= a.item_rows) a.idflag[2] = 1;
              myid = (int)min(max(v, 0LL), a.item_rows - 1);
              if (a.save) a.seq32[b * a.L + l0 + lane] = myid;
            }
            const int n = min(32, a.L - l0);
            // batches of 8 independent row loads in flight, accumulated in sequence order (ref :172)
            for (int l = 0; l < n; l += 8) {
              int id[8];
              float4 r[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                id[u] = __shfl_sync(0xffffffffu, myid, min(l + u, 31));
                if (l + u >= n) id[u] = 0;
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) r[u] = id[u] ? ld4(item_row<SHARDED>(a, id[u]) + 4 * lane) : f4(0.f);
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (id[u]) { acc += r[u]; ++nvalid; }
            }
          }
          cntv[s] = (float)max(nvalid, 1);           // clamp(min=1), ref :173
          f_hist[s] = acc / cntv[s];                  // ref :174
        }
        if (a.save && lane == 0) {
          int4 rec = make_int4((int)iid, (int)lk, (int)vw, nvalid);
          *reinterpret_cast<int4*>(a.ids + b * 4) = rec;
          a.cnt[b] = cntv[s];
        }
      }
      if (!EXTP) {   // stage the multimodal vector k-major: xs[k][s]
        xs[(4 * lane + 0) * EMB_SPW + s] = xm.x;
        xs[(4 * lane + 1) * EMB_SPW + s] = xm.y;
        xs[(4 * lane + 2) * EMB_SPW + s] = xm.z;
        xs[(4 * lane + 3) * EMB_SPW + s] = xm.w;
      }
    }
    __syncwarp();
    // projection: y[s][4*lane + c] = sum_k x[s][k] * W[4*lane + c][k]   (ref :106)
    float4 y[EMB_SPW];
#pragma unroll
    for (int s = 0; s < EMB_SPW; ++s) {
      const long long b = g * EMB_SPW + s;
      y[s] = (EXTP && b < a.B) ? ld4s(a.yproj + b * D + 4 * lane) : f4(0.f);      // EXTP: x W^T + b from the GEMM
    }
#pragma unroll 8
    for (int k = 0; k < (EXTP ? 0 : D); ++k) {
      float xv[EMB_SPW];
      if (EMB_SPW == 2) {
        const float2 t = *reinterpret_cast<const float2*>(xs + k * EMB_SPW);
        xv[0] = t.x; xv[EMB_SPW - 1] = t.y;
      } else if (EMB_SPW == 4) {
        const float4 t = *reinterpret_cast<const float4*>(xs + k * EMB_SPW);
        xv[0] = t.x; xv[1 % EMB_SPW] = t.y; xv[2 % EMB_SPW] = t.z; xv[3 % EMB_SPW] = t.w;
      } else {
#pragma unroll
        for (int s = 0; s < EMB_SPW; ++s) xv[s] = xs[k * EMB_SPW + s];
      }
      const float4 wv = *reinterpret_cast<const float4*>(Wt + k * D + 4 * (lane ^ (k & 31)));
#pragma unroll
      for (int s = 0; s < EMB_SPW; ++s) {
        y[s].x = fmaf(xv[s], wv.x, y[s].x); y[s].y = fmaf(xv[s], wv.y, y[s].y);
        y[s].z = fmaf(xv[s], wv.z, y[s].z); y[s].w = fmaf(xv[s], wv.w, y[s].w);
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < EMB_SPW; ++s) {
      const long long b = g * EMB_SPW + s;
      if (b >= a.B) continue;  // warp-uniform
      // LayerNorm(eps 1e-5, biased var) + ReLU  (ref :107-108)
      float4 yv = y[s] + bias;
      const float mean = warp_sum(hsum4(yv)) * (1.0f / D);
      const float4 dv = yv - f4(mean);
      const float var = warp_sum(hsum4(dv * dv)) * (1.0f / D);
      const float rs = 1.0f / sqrtf(var + 1e-5f);
      const float4 xh = dv * rs;
      float4 img = xh * gam + bet;
      img.x = fmaxf(img.x, 0.f); img.y = fmaxf(img.y, 0.f); img.z = fmaxf(img.z, 0.f); img.w = fmaxf(img.w, 0.f);
      // SENET (ref :28-35): z_f = mean_d x_f ; field 0 is the zero vector
      float z[NF];
      z[0] = 0.f;
      z[1] = warp_sum(hsum4(f_like[s])) * (1.0f / D);
      z[2] = warp_sum(hsum4(f_view[s])) * (1.0f / D);
      z[3] = warp_sum(hsum4(f_item[s])) * (1.0f / D);
      z[4] = warp_sum(hsum4(img)) * (1.0f / D);
      z[5] = warp_sum(hsum4(f_hist[s])) * (1.0f / D);
      float h[SE_R];
#pragma unroll
      for (int r = 0; r < SE_R; ++r) {
        float acc = b1[r];
#pragma unroll
        for (int f = 0; f < NF; ++f) acc = fmaf(z[f], w1[r * NF + f], acc);
        h[r] = fmaxf(acc, 0.f);
      }
      float sg[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        float acc = b2[f];
#pragma unroll
        for (int r = 0; r < SE_R; ++r) acc = fmaf(h[r], w2[f * SE_R + r], acc);
        sg[f] = sigmoidf_(acc);
      }
      float* crow = a.C + b * K1 + 4 * lane;
      const float4 v1 = f_like[s] * sg[1], v2 = f_view[s] * sg[2], v3 = f_item[s] * sg[3], v4 = img * sg[4], v5 = f_hist[s] * sg[5];
      st4(crow + 1 * D, v1);
      st4(crow + 2 * D, v2);
      st4(crow + 3 * D, v3);
      st4(crow + 4 * D, v4);
      st4(crow + 5 * D, v5);
      if (a.pkC.mode) {     // the GEMMs read this copy (tf32 hi|lo or bf16): no separate pack pass over C
        const long long e = b * K1 + 4 * lane;
        store_packed4(a.pkC.base, a.pkC.lo_off, a.pkC.mode, e + 1 * D, v1);
        store_packed4(a.pkC.base, a.pkC.lo_off, a.pkC.mode, e + 2 * D, v2);
        store_packed4(a.pkC.base, a.pkC.lo_off, a.pkC.mode, e + 3 * D, v3);
        store_packed4(a.pkC.base, a.pkC.lo_off, a.pkC.mode, e + 4 * D, v4);
        store_packed4(a.pkC.base, a.pkC.lo_off, a.pkC.mode, e + 5 * D, v5);
      }
      if (a.save) {
        float* xrow = a.X5 + b * (NA * D) + 4 * lane;
        st4(xrow + 0 * D, f_like[s]);
        st4(xrow + 1 * D, f_view[s]);
        st4(xrow + 2 * D, f_item[s]);
        st4(xrow + 3 * D, img);
        st4(xrow + 4 * D, f_hist[s]);
        st4(a.xhat + b * D + 4 * lane, xh);
        {
          float v = 0.f;
#pragma unroll
          for (int f = 0; f < NF; ++f) if (lane == f) v = sg[f];
          if (lane < 8) a.sgate[b * 8 + lane] = v;
        }
        if (lane == 0) a.rstd[b] = rs;
      }
    }
  }
}

size_t embed_fwd_smem(bool extp) { return extp ? 0 : (size_t)(D * D + EMB_WARPS * D * EMB_SPW) * sizeof(float); }

template <int SE_R, bool EXTP>
static int launch_fwd_r(const EmbedFwdArgs& a, unsigned blocks, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set && smem > 0) {
    FBN_CHECK_CUDA(cudaFuncSetAttribute(embed_senet_fwd_kernel<false, SE_R, EXTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FBN_CHECK_CUDA(cudaFuncSetAttribute(embed_senet_fwd_kernel<true, SE_R, EXTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (a.nshard > 0) embed_senet_fwd_kernel<true, SE_R, EXTP><<<blocks, EMB_WARPS * 32, smem, st>>>(a);
  else embed_senet_fwd_kernel<false, SE_R, EXTP><<<blocks, EMB_WARPS * 32, smem, st>>>(a);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// SENET hidden width = max(1, 6 // reduction_ratio) (ref :13): 3 for the reference's ratio 2; 6 / 2 / 1 for ratios 1 / 3 / >= 4
int launch_embed_senet_fwd(const EmbedFwdArgs& a, cudaStream_t st) {
  const bool extp = a.yproj != nullptr;
  const size_t smem = embed_fwd_smem(extp);
  const long long ngroups = (a.B + EMB_SPW - 1) / EMB_SPW;
  long long blocks = (ngroups + EMB_WARPS - 1) / EMB_WARPS;
  // persistent: 2 resident CTAs per SM with the in-kernel projection (72 KB smem each; 3 per SM measured slower: 342 vs 320 us),
  // 3 without it
  const long long cap = (extp ? 3LL : 2LL) * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  switch (a.se_r * 2 + (extp ? 1 : 0)) {
    case 2: return launch_fwd_r<1, false>(a, (unsigned)blocks, smem, st);
    case 3: return launch_fwd_r<1, true>(a, (unsigned)blocks, smem, st);
    case 4: return launch_fwd_r<2, false>(a, (unsigned)blocks, smem, st);
    case 5: return launch_fwd_r<2, true>(a, (unsigned)blocks, smem, st);
    case 6: return launch_fwd_r<3, false>(a, (unsigned)blocks, smem, st);
    case 7: return launch_fwd_r<3, true>(a, (unsigned)blocks, smem, st);
    case 12: return launch_fwd_r<6, false>(a, (unsigned)blocks, smem, st);
    case 13: return launch_fwd_r<6, true>(a, (unsigned)blocks, smem, st);
  }
  FBN_REQUIRE(false, FBN_ERR_SHAPE, "SENET hidden width %d is not one of 1, 2, 3, 6 (= max(1, 6 // reduction_ratio))", a.se_r);
}

// ---------------------------------------------------------------------------------------------
// Backward of SENET + field stack + LayerNorm/ReLU + pooling (everything up to the table rows).
// ---------------------------------------------------------------------------------------------

constexpr int EBW_WARPS = 8;

template <int SE_R>
__global__ void __launch_bounds__(EBW_WARPS * 32) embed_senet_bwd_kernel(EmbedBwdArgs a) {
  extern __shared__ __align__(16) float smem[];  // [EBW_WARPS][cate_rows][128] private accumulators
  __shared__ float s_se[SE_R * NF + SE_R + NF * SE_R];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < SE_R * NF) s_se[threadIdx.x] = a.se_w1[threadIdx.x];
  if (threadIdx.x < SE_R) s_se[SE_R * NF + threadIdx.x] = a.se_b1[threadIdx.x];
  if (threadIdx.x < NF * SE_R) s_se[SE_R * NF + SE_R + threadIdx.x] = a.se_w2[threadIdx.x];
  float* acc = smem + warp * (a.cate_rows * D);
  for (int i = lane; i < a.cate_rows * D; i += 32) acc[i] = 0.f;
  __syncthreads();
  const float* w1 = s_se; const float* b1 = s_se + SE_R * NF; const float* w2 = b1 + SE_R;
  const float4 gam = ld4(a.ln_g + 4 * lane);

  // contiguous sample range per warp so that the accumulation order is fixed by (grid, B) only
  const long long nw = (long long)gridDim.x * EBW_WARPS;
  const long long per = (a.B + nw - 1) / nw;
  const long long wid = (long long)blockIdx.x * EBW_WARPS + warp;
  const long long b0 = wid * per, b1e = min(a.B, b0 + per);
  for (long long b = b0; b < b1e; ++b) {
    float4 x[NA], dv[NA];
#pragma unroll
    for (int f = 0; f < NA; ++f) {
      x[f] = ld4s(a.X5 + b * (NA * D) + f * D + 4 * lane);
      dv[f] = ld4s(a.dV + b * (NA * D) + f * D + 4 * lane);
    }
    float sg[NF];
    {
      float t = lane < 8 ? a.sgate[b * 8 + lane] : 0.f;
#pragma unroll
      for (int f = 0; f < NF; ++f) sg[f] = __shfl_sync(0xffffffffu, t, f);
    }
    float z[NF], ds[NF];
    z[0] = 0.f; ds[0] = 0.f;
#pragma unroll
    for (int f = 0; f < NA; ++f) {
      z[f + 1] = warp_sum(hsum4(x[f])) * (1.0f / D);
      ds[f + 1] = warp_sum(hsum4(dv[f] * x[f]));
    }
    float h[SE_R];
#pragma unroll
    for (int r = 0; r < SE_R; ++r) {
      float t = b1[r];
#pragma unroll
      for (int f = 0; f < NF; ++f) t = fmaf(z[f], w1[r * NF + f], t);
      h[r] = fmaxf(t, 0.f);
    }
    float da2[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) da2[f] = ds[f] * sg[f] * (1.0f - sg[f]);
    float da1[SE_R];
#pragma unroll
    for (int r = 0; r < SE_R; ++r) {
      float t = 0.f;
#pragma unroll
      for (int f = 0; f < NF; ++f) t = fmaf(da2[f], w2[f * SE_R + r], t);
      da1[r] = h[r] > 0.f ? t : 0.f;
    }
    float4 dx[NA];
#pragma unroll
    for (int f = 0; f < NA; ++f) {
      float dz = 0.f;
#pragma unroll
      for (int r = 0; r < SE_R; ++r) dz = fmaf(da1[r], w1[r * NF + f + 1], dz);
      dx[f] = dv[f] * sg[f + 1] + f4(dz * (1.0f / D));
    }
    {
      float v = 0.f;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        if (lane == f) v = da2[f];
        if (lane == NF + SE_R + f) v = z[f];
      }
#pragma unroll
      for (int r = 0; r < SE_R; ++r) {
        if (lane == NF + r) v = h[r];
        if (lane == 2 * NF + SE_R + r) v = da1[r];
      }
      if (lane < 24) a.sestat[b * 24 + lane] = v;      // record {da2[6], h[R], z[6], da1[R]}: 12 + 2R <= 24 floats
    }
    const int4 id = *reinterpret_cast<const int4*>(a.ids + b * 4);
    // fields 1,2 -> cate_emb rows (private per-warp accumulators: deterministic, no atomics)
    {
      float4* p = reinterpret_cast<float4*>(acc + id.y * D + 4 * lane);
      *p = *p + dx[0];
      __syncwarp();
      float4* q = reinterpret_cast<float4*>(acc + id.z * D + 4 * lane);
      *q = *q + dx[1];
      __syncwarp();
    }
    st4(a.dXitem + b * D + 4 * lane, dx[2]);                               // field 3 -> item_emb[item_id]
    st4(a.dXhist + b * D + 4 * lane, dx[4] / a.cnt[b]);                     // field 5 -> each history row
    // field 4: ReLU -> LayerNorm backward
    float4 dl = dx[3];
    dl.x = x[3].x > 0.f ? dl.x : 0.f; dl.y = x[3].y > 0.f ? dl.y : 0.f;
    dl.z = x[3].z > 0.f ? dl.z : 0.f; dl.w = x[3].w > 0.f ? dl.w : 0.f;
    st4(a.dln + b * D + 4 * lane, dl);
    const float4 xh = ld4s(a.xhat + b * D + 4 * lane);
    const float4 dxh = dl * gam;
    const float m1 = warp_sum(hsum4(dxh)) * (1.0f / D);
    const float m2 = warp_sum(hsum4(dxh * xh)) * (1.0f / D);
    const float rs = a.rstd[b];
    const float4 dyv = (dxh - f4(m1) - xh * m2) * rs;
    st4(a.dy + b * D + 4 * lane, dyv);
    if (a.pkdy.mode) store_packed4(a.pkdy.base, a.pkdy.lo_off, a.pkdy.mode, b * D + 4 * lane, dyv);
  }
  __syncthreads();
  // fixed-order reduction of the warps' private accumulators -> one partial per CTA
  for (int i = threadIdx.x; i < a.cate_rows * D; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < EBW_WARPS; ++w) t += smem[w * (a.cate_rows * D) + i];
    a.cate_partial[(long long)blockIdx.x * (a.cate_rows * D) + i] = t;
  }
}

int embed_bwd_blocks(long long B) {
  // >= 4 samples per warp: at the reference's batch 4096 that is 128 CTAs (16 per warp left 116 of the 148 SMs idle: 43 us)
  long long blocks = (B + EBW_WARPS * 4 - 1) / (EBW_WARPS * 4);
  long long cap = 2LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int launch_embed_senet_bwd(const EmbedBwdArgs& a, int blocks, cudaStream_t st) {
  const size_t smem = (size_t)EBW_WARPS * a.cate_rows * D * sizeof(float);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    FBN_CHECK_CUDA(cudaFuncSetAttribute(embed_senet_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FBN_CHECK_CUDA(cudaFuncSetAttribute(embed_senet_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FBN_CHECK_CUDA(cudaFuncSetAttribute(embed_senet_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FBN_CHECK_CUDA(cudaFuncSetAttribute(embed_senet_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  switch (a.se_r) {
    case 1: embed_senet_bwd_kernel<1><<<blocks, EBW_WARPS * 32, smem, st>>>(a); break;
    case 2: embed_senet_bwd_kernel<2><<<blocks, EBW_WARPS * 32, smem, st>>>(a); break;
    case 3: embed_senet_bwd_kernel<3><<<blocks, EBW_WARPS * 32, smem, st>>>(a); break;
    case 6: embed_senet_bwd_kernel<6><<<blocks, EBW_WARPS * 32, smem, st>>>(a); break;
    default: FBN_REQUIRE(false, FBN_ERR_SHAPE, "SENET hidden width %d is not one of 1, 2, 3, 6", a.se_r);
  }
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// sums `parts` consecutive partial vectors of length n in index order: out[i] = sum_p partial[p][i]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int parts, long long n,
                                       int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int p = 0; p < parts; ++p) t += partial[(long long)p * n + i];
    out[i] = accumulate ? out[i] + t : t;
  }
}

int launch_reduce_partials(const float* partial, float* out, int parts, long long n, int accumulate, cudaStream_t st) {
  int blocks = (int)std::min<long long>((n + 255) / 256, 4LL * num_sms());
  reduce_partials_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(partial, out, parts, n, accumulate);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// SENET parameter gradients from the per-sample records written above (R = hidden width):
//  sestat[b] = {da2[0..5], h[0..R-1], z[0..5], da1[0..R-1], pad}
//  dW2[f][r] = sum_b da2[f] h[r]; db2[f] = sum_b da2[f]; dW1[r][f] = sum_b da1[r] z[f]; db1[r] = sum_b da1[r]
// output order (13R + 6 values): dW1 (6R), db1 (R), dW2 (6R), db2 (6)
constexpr int SE_PSTRIDE = 96;     // >= 13 * 6 + 6
__global__ void senet_param_partial_kernel(const float* __restrict__ sestat, long long B, long long per, int R, float* __restrict__ partial) {
  const int t = threadIdx.x, nout = 13 * R + NF;
  if (t >= nout) return;
  int ia, ib;  // indices into the 24-float record; ib = -1 -> times 1
  if (t < NF * R) { ia = 2 * NF + R + t / NF; ib = NF + R + t % NF; }
  else if (t < NF * R + R) { ia = 2 * NF + R + (t - NF * R); ib = -1; }
  else if (t < 2 * NF * R + R) { ia = (t - NF * R - R) / R; ib = NF + (t - NF * R - R) % R; }
  else { ia = t - 2 * NF * R - R; ib = -1; }
  const long long b0 = (long long)blockIdx.x * per, b1 = min(B, b0 + per);
  float acc = 0.f;
  for (long long b = b0; b < b1; ++b) {
    const float* r = sestat + b * 24;
    acc += ib >= 0 ? r[ia] * r[ib] : r[ia];
  }
  partial[(long long)blockIdx.x * SE_PSTRIDE + t] = acc;
}

__global__ void __launch_bounds__(256) senet_param_final_kernel(const float* __restrict__ partial, int parts, int R, float* dw1, float* db1,
                                                                float* dw2, float* db2) {
  const int lane = threadIdx.x & 31, nout = 13 * R + NF;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);   // warp per output, lanes over the partials (fixed xor tree)
  if (t >= nout) return;
  double acc = 0.0;
  for (int p = lane; p < parts; p += 32) acc += (double)partial[(long long)p * SE_PSTRIDE + t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane != 0) return;
  if (t < NF * R) dw1[t] = (float)acc;
  else if (t < NF * R + R) db1[t - NF * R] = (float)acc;
  else if (t < 2 * NF * R + R) dw2[t - NF * R - R] = (float)acc;
  else db2[t - 2 * NF * R - R] = (float)acc;
}

int launch_senet_param_grads(const float* sestat, long long B, int R, float* partial, float* dw1, float* db1, float* dw2, float* db2,
                             cudaStream_t st) {
  int parts = (int)std::min<long long>((B + 31) / 32, 1024);
  if (parts < 1) parts = 1;
  long long per = (B + parts - 1) / parts;
  senet_param_partial_kernel<<<parts, 96, 0, st>>>(sestat, B, per, R, partial);
  FBN_CHECK_LAUNCH();
  senet_param_final_kernel<<<(13 * R + NF + 7) / 8, 256, 0, st>>>(partial, parts, R, dw1, db1, dw2, db2);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

}  // namespace fbn

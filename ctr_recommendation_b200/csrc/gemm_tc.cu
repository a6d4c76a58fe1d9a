// tcgen05 / TMA / TMEM GEMM back end for the dense contractions (bilinear + DNN tower).
//
//   FBN_PREC_TF32X3 : fp32-grade result from three kind::tf32 passes  D = Ah*Bh + Ah*Bl + Al*Bh  with
//                     x_h = x & 0xffffe000 (exactly representable in tf32) and x_l = x - x_h, fp32 accumulation
//                     in TMEM (single-pass TF32 misses the 1e-5 logit tolerance, 3x passes it).
//   FBN_PREC_BF16   : one kind::f16 (bf16) pass, fp32 accumulation in TMEM.
//
// Structure (one 128x128 output tile per CTA, 192 threads):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor.2d (128B swizzle) into a 3..6 stage smem ring
//   warp 1      : MMA issuer    -- one elected thread issues tcgen05.mma, tcgen05.commit frees the stage
//   warps 2..5  : epilogue      -- tcgen05.ld the fp32 accumulators (TMEM lane quarter = warp % 4),
//                                  add bias / accumulate, store fp32 rows
// Accumulation is CHUNKED: the tensor core adds into its fp32 TMEM accumulator with truncation, which biases a
// long K loop (measured 2.2e-5 relative at K = 2688 with 3xTF32 -- no better than one TF32 pass).  So the MMA warp
// switches between two TMEM accumulators every TC_CHUNK k-blocks and the epilogue warps fold each finished chunk
// into fp32 registers with round-to-nearest adds while the next chunk is being computed.
// Both operands are consumed K-major.  A "pack" pre-pass (pack kernels below) converts the fp32 activations /
// weights to the operand format (hi|lo split or bf16) and transposes when the contraction runs over the
// leading dimension (weight gradients), so every GEMM flavour of the step maps onto this one kernel.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm.h"

namespace fbn {

constexpr int TC_BM = 128, TC_BN = 128;
constexpr int TC_THREADS = 192;
constexpr int TC_CHUNK = 4;   // k-blocks accumulated inside TMEM before draining to registers

struct TcArgs {
  float* C; const float* bias;
  long long M, N, K, ldc;
  int splits; long long strideSplit;
  unsigned long long kmask, nmask;
  int accumulate;
  int a_lo_row, b_lo_row;  // row offset of the "lo" half inside the packed operand (tf32x3)
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, 128-byte swizzle shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start >> 4 ; [16,30) LBO >> 4 (=1, unused for swizzled K-major) ; [32,46) SBO >> 4 = 1024 B between
//   8-row groups ; [46,48) version = 1 ; [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6); a/b format @ [7,10) / [10,13)
// (BF16 = 1, TF32 = 2); K-major A and B; N >> 3 @ [17,23); M >> 4 @ [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int MODE>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (MODE == FBN_PREC_TF32X3) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct TcCfg {
  static constexpr int BK = MODE == FBN_PREC_TF32X3 ? 32 : 64;          // elements per 128-byte swizzle row
  static constexpr int UK = MODE == FBN_PREC_TF32X3 ? 8 : 16;           // K per tcgen05.mma
  static constexpr int NPART = MODE == FBN_PREC_TF32X3 ? 2 : 1;         // hi + lo
  static constexpr int TILE_BYTES = TC_BM * 128;                        // one 128-row x 128-byte operand tile
  static constexpr int STAGE_BYTES = 2 * NPART * TILE_BYTES;            // A parts + B parts
  static constexpr int STAGES = MODE == FBN_PREC_TF32X3 ? 3 : 6;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int NBAR = 2 * STAGES + 4;
  static constexpr int FMT = MODE == FBN_PREC_TF32X3 ? 2 : 1;
};

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs g) {
  using Cfg = TcCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;        // [2] accumulator buffer complete
  uint64_t* tempty = bars + 2 * Cfg::STAGES + 2;   // [2] accumulator buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM, sp = blockIdx.z;
  if (g.nmask != ~0ull && !((g.nmask >> (n0 / 128)) & 1ull)) return;

  // K blocks of this split (contiguous range); blocks inside structurally-zero 128-column groups are skipped
  const int kblocks = (int)((g.K + Cfg::BK - 1) / Cfg::BK);
  const int per = (kblocks + g.splits - 1) / g.splits;
  const int kb0 = sp * per, kb1 = min(kblocks, kb0 + per);
  auto active = [&](int kb) { return g.kmask == ~0ull || ((g.kmask >> ((kb * Cfg::BK) / 128)) & 1ull); };
  int nact = 0;
  for (int kb = kb0; kb < kb1; ++kb) nact += active(kb) ? 1 : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM: two accumulators of 128 lanes x 128 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nact > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!active(kb)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, Cfg::STAGE_BYTES);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        const int kc = kb * Cfg::BK;
        tma_load_2d(st, &tmA, full + s, kc, m0);
        tma_load_2d(st + Cfg::NPART * Cfg::TILE_BYTES, &tmB, full + s, kc, n0);
        if (Cfg::NPART == 2) {
          tma_load_2d(st + Cfg::TILE_BYTES, &tmA, full + s, kc, g.a_lo_row + m0);
          tma_load_2d(st + 3 * Cfg::TILE_BYTES, &tmB, full + s, kc, g.b_lo_row + n0);
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nact > 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::FMT, TC_BM, TC_BN);
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!active(kb)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        const int chunk = it / TC_CHUNK, buf = chunk & 1, pos = it % TC_CHUNK;
        if (pos == 0) {  // the epilogue must have drained this accumulator (two chunks ago)
          mbar_wait(tempty + buf, ((chunk >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * TC_BN);
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t a_hi = make_desc(sa), b_hi = make_desc(sa + Cfg::NPART * Cfg::TILE_BYTES);
#pragma unroll
        for (int k = 0; k < Cfg::BK / Cfg::UK; ++k) {
          const uint64_t adv = (uint64_t)((k * Cfg::UK * (MODE == FBN_PREC_TF32X3 ? 4 : 2)) >> 4);
          const uint32_t acc = (pos > 0 || k > 0) ? 1u : 0u;
          if (Cfg::NPART == 2) {
            const uint64_t a_lo = make_desc(sa + Cfg::TILE_BYTES), b_lo = make_desc(sa + 3 * Cfg::TILE_BYTES);
            umma<MODE>(tacc, a_lo + adv, b_hi + adv, idesc, acc);      // small terms first
            umma<MODE>(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
            umma<MODE>(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
          } else {
            umma<MODE>(tacc, a_hi + adv, b_hi + adv, idesc, acc);
          }
        }
        tc_commit(empty + s);                                  // frees the smem stage once the MMAs have read it
        if (pos == TC_CHUNK - 1 || it == nact - 1) tc_commit(tfull + buf);   // chunk complete
        ++it;
      }
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 == output rows m0 + 32*(w%4) + lane
    const int q = warp & 3;
    const long long row = (long long)m0 + q * 32 + lane;
    float acc[TC_BN];
#pragma unroll
    for (int j = 0; j < TC_BN; ++j) acc[j] = 0.f;
    const int nchunks = (nact + TC_CHUNK - 1) / TC_CHUNK;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(tfull + buf, (c >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < TC_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_BN + c0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);   // round-to-nearest fold of the chunk
      }
      tc_fence_before();
      mbar_arrive(tempty + buf);
    }
    if (row < g.M) {
      float* crow = g.C + (long long)sp * g.strideSplit + row * g.ldc + n0;
#pragma unroll
      for (int j = 0; j < TC_BN; j += 4) {
        float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        if (g.bias) o += ld4(g.bias + n0 + j);
        if (g.accumulate) o += *reinterpret_cast<const float4*>(crow + j);
        st4(crow + j, o);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// pack kernels: fp32 (rows x cols, ld) -> K-major operand, optionally transposed
//   tf32x3: dst[0..R) = hi, dst[lo_row .. lo_row+R) = lo, row pitch Kp floats
//   bf16  : dst (R x Kp) bf16
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <int MODE>
__global__ void pack_rows_kernel(const float* __restrict__ src, long long ld, long long R, long long K, long long Kp, void* dst,
                                 long long lo_row) {
  const long long q = Kp / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * q; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / q, c = (i % q) * 4;
    float4 v = f4(0.f);
    if (c + 3 < K) v = ld4s(src + r * ld + c);
    else {
      if (c < K) v.x = src[r * ld + c];
      if (c + 1 < K) v.y = src[r * ld + c + 1];
      if (c + 2 < K) v.z = src[r * ld + c + 2];
    }
    if (MODE == FBN_PREC_TF32X3) {
      float* d = static_cast<float*>(dst);
      const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      st4(d + r * Kp + c, h);
      st4(d + (lo_row + r) * Kp + c, v - h);
    } else {
      __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst);
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&p0);
      o.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(d + r * Kp + c) = o;
    }
  }
}

// src stored (K rows, R cols) -> dst[r][k] = src[k][r]
template <int MODE>
__global__ void __launch_bounds__(256) pack_transpose_kernel(const float* __restrict__ src, long long ld, long long R, long long K,
                                                             long long Kp, void* dst, long long lo_row) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long long r0 = (long long)blockIdx.x * 32, k0 = (long long)blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long k = k0 + ty + i * 8, r = r0 + tx;
    tile[ty + i * 8][tx] = (k < K && r < R) ? src[k * ld + r] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + ty + i * 8, k = k0 + tx;
    if (r < R && k < Kp) {
      const float v = tile[tx][ty + i * 8];
      if (MODE == FBN_PREC_TF32X3) {
        float* d = static_cast<float*>(dst);
        const float h = tf32_hi(v);
        d[r * Kp + k] = h;
        d[(lo_row + r) * Kp + k] = v - h;
      } else {
        static_cast<__nv_bfloat16*>(dst)[r * Kp + k] = __float2bfloat16_rn(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D map over a (rows, K) K-contiguous operand; box = (128 bytes of K) x 128 rows, 128B swizzle, OOB -> 0
static int make_map(CUtensorMap* m, int mode, void* base, long long rows, long long K, long long pitch_elems) {
  EncodeTiledFn enc = get_encode();
  FBN_REQUIRE(enc != nullptr, FBN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int esz = mode == FBN_PREC_TF32X3 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(pitch_elems * esz)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, mode == FBN_PREC_TF32X3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FBN_REQUIRE(r == CUDA_SUCCESS, FBN_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows %lld K %lld pitch %lld)", (int)r, rows, K,
              pitch_elems);
  return FBN_OK;
}

static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// bytes of scratch one gemm_tc call needs for the packed operands
size_t gemm_tc_scratch_bytes(long long M, long long N, long long K, int precision) {
  const long long Kp = round_up(K, 8);
  const long long parts = precision == FBN_PREC_TF32X3 ? 2 : 1;
  const long long esz = precision == FBN_PREC_TF32X3 ? 4 : 2;
  const long long Mp = round_up(M, 128), Np = round_up(N, 128);
  return (size_t)((Mp + Np) * parts * Kp * esz + 2048);
}

bool gemm_tc_supported(const GemmArgs& g, int precision) {
  return (precision == FBN_PREC_TF32X3 || precision == FBN_PREC_BF16) && g.N % 128 == 0 && g.ldc % 4 == 0 && g.batch >= 1;
}

template <int MODE>
static int pack_operand(const float* src, long long ld, bool stored_k_major, long long R, long long K, long long Kp, void* dst,
                        long long lo_row, cudaStream_t st) {
  if (stored_k_major) {  // (R, K) row-major: straight conversion
    const long long n = R * (Kp / 4);
    int blocks = (int)std::min<long long>(cdiv(n, 256), 16LL * num_sms());
    pack_rows_kernel<MODE><<<std::max(blocks, 1), 256, 0, st>>>(src, ld, R, K, Kp, dst, lo_row);
  } else {               // stored (K, R): transpose
    dim3 grid((unsigned)cdiv(R, 32), (unsigned)cdiv(Kp, 32));
    pack_transpose_kernel<MODE><<<grid, 256, 0, st>>>(src, ld, R, K, Kp, dst, lo_row);
  }
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

template <int MODE>
static int gemm_tc_mode(const GemmArgs& g, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  using Cfg = TcCfg<MODE>;
  static bool attr = false;
  if (!attr) {
    FBN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const long long Kp = round_up(g.K, 8);
  const long long Mp = round_up(g.M, 128), Np = round_up(g.N, 128);
  const long long esz = MODE == FBN_PREC_TF32X3 ? 4 : 2;
  const size_t a_bytes = (size_t)(Mp * Cfg::NPART * Kp * esz), b_bytes = (size_t)(Np * Cfg::NPART * Kp * esz);
  FBN_REQUIRE(scratch != nullptr && scratch_bytes >= a_bytes + b_bytes + 2048, FBN_ERR_ARG,
              "tcgen05 GEMM needs %zu bytes of operand scratch, got %zu", a_bytes + b_bytes + 2048, scratch_bytes);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~uintptr_t(1023));
  void* packA = base;
  void* packB = base + ((a_bytes + 1023) & ~size_t(1023));
  for (int bi = 0; bi < g.batch; ++bi) {
    const float* A = g.A + bi * g.strideA;
    const float* Bm = g.B + bi * g.strideB;
    // A: a_t == 0 -> stored (M,K) = K-major ; a_t != 0 -> stored (K,M)
    int rc = pack_operand<MODE>(A, g.lda, g.a_t == 0, g.M, g.K, Kp, packA, Mp, st);
    if (rc) return rc;
    // B: b_t != 0 -> stored (N,K) = K-major ; b_t == 0 -> stored (K,N)
    rc = pack_operand<MODE>(Bm, g.ldb, g.b_t != 0, g.N, g.K, Kp, packB, Np, st);
    if (rc) return rc;
    CUtensorMap tmA, tmB;
    rc = make_map(&tmA, MODE, packA, Mp * Cfg::NPART, g.K, Kp);
    if (rc) return rc;
    rc = make_map(&tmB, MODE, packB, Np * Cfg::NPART, g.K, Kp);
    if (rc) return rc;
    TcArgs t;
    t.C = g.C + bi * g.strideC; t.bias = g.bias; t.M = g.M; t.N = g.N; t.K = g.K; t.ldc = g.ldc; t.splits = g.splits;
    t.strideSplit = g.strideSplit; t.kmask = g.kmask; t.nmask = g.nmask; t.accumulate = g.accumulate;
    t.a_lo_row = (int)Mp; t.b_lo_row = (int)Np;
    dim3 grid((unsigned)(g.N / TC_BN), (unsigned)cdiv(g.M, TC_BM), (unsigned)g.splits);
    gemm_tc_kernel<MODE><<<grid, TC_THREADS, Cfg::SMEM, st>>>(tmA, tmB, t);
    FBN_CHECK_LAUNCH();
  }
  return FBN_OK;
}

int gemm_tc(const GemmArgs& g, int precision, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  FBN_REQUIRE(gemm_tc_supported(g, precision), FBN_ERR_SHAPE, "tcgen05 GEMM: unsupported shape (N %lld must be a multiple of 128)", g.N);
  if (g.M <= 0) return FBN_OK;
  if (precision == FBN_PREC_TF32X3) return gemm_tc_mode<FBN_PREC_TF32X3>(g, scratch, scratch_bytes, st);
  return gemm_tc_mode<FBN_PREC_BF16>(g, scratch, scratch_bytes, st);
}

}  // namespace fbn

// tcgen05 / TMA / TMEM GEMM back end for the dense contractions (bilinear + DNN tower).
//
//   FBN_PREC_TF32X3 : fp32-grade result from three kind::tf32 passes  D = Ah*Bh + Ah*Bl + Al*Bh  with
//                     x_h = x & 0xffffe000 (exactly representable in tf32) and x_l = x - x_h, fp32 accumulation
//                     in TMEM (single-pass TF32 misses the 1e-5 logit tolerance, 3x passes it).
//   FBN_PREC_BF16   : one kind::f16 (bf16) pass, fp32 accumulation in TMEM.
//   FBN_PREC_TF32X2 : Ah*Bh as kind::tf32 plus Al*Bh + Ah*Bl as kind::f16 bf16 MMAs into the same accumulator.  The correction
//                     terms are 2^-11 of the product, so bf16 operands (2^-9) keep the result at ~1.4e-6 while the tensor work
//                     drops from 3 to 1 + 1/2 + 1/2 = 2 pass equivalents.  Operand part 1 holds, per 32-element k-block, one
//                     128-byte row [hi as bf16 x32 | lo as bf16 x32]: the same 128B-swizzled K-major tile as the fp32 lo part,
//                     the bf16 MMAs address the two halves by the ordinary in-row descriptor advance.  K-major operands only.
//
//   FBN_PREC_F16X3  : fp32-grade result from three kind::f16 (fp16) passes at the bf16 rate -- half the tensor time AND half the
//                     operand bytes of tf32x3.  Each operand tensor carries ONE power-of-two scale s (amax * s in [2^14, 2^15)):
//                     x_h = fp16_rn(s x), x_l = fp16_rn(s x - x_h), i.e. the same 22 operand bits as the tf32 split (fp16 and tf32
//                     both keep 11 significand bits); D = (Ah*Bh + Ah*Bl + Al*Bh) / (sa sb), the division being two exact
//                     multiplications by powers of two in the epilogue.  Elements more than 2^28 below the tensor's amax lose
//                     relative (not absolute) precision: their error is <= 2^-25 / s, i.e. <= 2^-39 of amax
//                     (tools/split_precision_sim.py: logits 7.6e-7, worst gradient 6e-6 -- the same as tf32x3 and as fp32 itself,
//                     and unchanged when the scale is off by a factor of 2^12).  The amax pass makes the scale exact and the
//                     result independent of anything but the operand values (no state carried between calls).
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
// Operands are "packed" once per step by an elementwise pre-pass into the MMA operand format (tf32 hi | lo split, or
// bf16) in their NATURAL row-major layout -- no transposes: an operand whose contraction index is the leading (row)
// index is consumed MN-major, the other way K-major, so every GEMM flavour of the step (forward NT, data-gradient NN,
// weight-gradient TN) maps onto this one kernel and a packed tensor is shared by all GEMMs that read it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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
  // batched launch (blockIdx.z = batch * splits + split): per-batch TMA coordinate offsets and output stride
  int a_bc = 0, a_br = 0, b_bc = 0, b_br = 0;
  long long strideC = 0;
  // FBN_PREC_F16X3: device pointers to the operands' inverse scales (1 / s, a power of two), applied in the epilogue
  const float* inv_sa = nullptr;
  const float* inv_sb = nullptr;
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
//
// MN-major operands (stored [K][MN], the natural layout of an activation whose batch dimension is contracted, or of
// an nn.Linear weight used transposed) use the canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)): one TMA box is
// [BK k-rows][128 bytes of MN]; 8-row K groups are SBO = 1024 B apart and successive 128-byte MN chunks are
// LBO = (bytes of one box) apart.
// 32-bit (tf32) MN-major operands only exist in the SWIZZLE_128B_BASE32B flavour (cutlass: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"): 32-byte chunks swizzled over 4-row groups (TMA mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), SBO = 512 B between the 4-row groups, layout type 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes = 16, uint32_t sbo_bytes = 1024, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6); a/b format @ [7,10) / [10,13)
// (BF16 = 1, TF32 = 2); a_major @ 15, b_major @ 16 (0 = K-major, 1 = MN-major); N >> 3 @ [17,23); M >> 4 @ [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr bool is32(int mode) { return mode == FBN_PREC_TF32X3 || mode == FBN_PREC_TF32X2; }   // 4-byte operand parts
constexpr bool two_part(int mode) { return is32(mode) || mode == FBN_PREC_F16X3; }             // hi + lo tiles: 64 KB per stage

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

// Epilogue store: each thread holds 128 consecutive fp32 of ONE output row (TMEM lane = row), which would make every
// store instruction of a warp touch 32 different rows (16 B each).  The warp instead transposes its 32 x 128 block through
// shared memory (the pipeline stages are idle by then) and writes full 512-byte row segments, one row per instruction.
constexpr int EPI_PITCH = 132;                      // floats per staged row (+4 pad)
constexpr int EPI_WARP_BYTES = 32 * EPI_PITCH * 4;  // 16.5 KB per epilogue warp

// shared-memory accesses of the staging buffer as explicit ld/st.shared: through a generic pointer the compiler emitted LD.E / ST.E
// and, not knowing that they cannot alias the global stores of the same loop, kept every staged read behind the previous row's
// store (ncu source view of the MLP-1 data gradient: 31 % of all stall samples sat on those generic loads)
__device__ __forceinline__ void sts128(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}

// s0 / s1: inverse operand scales of FBN_PREC_F16X3 (exact powers of two, applied one after the other so that their product can
// never leave the fp32 range on its own); 1 for the other modes.  stage: shared-space byte address of this warp's staging rows.
__device__ __forceinline__ void epilogue_store(const float (&acc)[128], uint32_t stage, int lane, long long row0, long long M,
                                               float* cbase, long long ldc, const float* bias, int accumulate, float s0 = 1.f,
                                               float s1 = 1.f) {
#pragma unroll
  for (int j = 0; j < 128; j += 4)
    sts128(stage + (uint32_t)(lane * EPI_PITCH + j) * 4u, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
  __syncwarp();
  const float4 bv = bias ? ld4(bias + 4 * lane) : f4(0.f);
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 8) {       // eight staged rows first, then their eight global stores
    if (row0 + r0 >= M) break;
    float4 o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = lds128(stage + (uint32_t)((r0 + i) * EPI_PITCH + 4 * lane) * 4u);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + i;
      if (row0 + r < M) {
        float4 x = (o[i] * s0) * s1 + bv;
        float* cp = cbase + (long long)r * ldc + 4 * lane;
        if (accumulate) x += *reinterpret_cast<const float4*>(cp);
        st4(cp, x);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct TcCfg {
  static constexpr int ESZ = is32(MODE) ? 4 : 2;
  static constexpr int BK = 128 / ESZ;                                   // k-block: 32 (tf32) / 64 (bf16) elements
  static constexpr int UK = 32 / ESZ;                                    // K per tcgen05.mma: 8 / 16
  static constexpr int EPB = 128 / ESZ;                                  // elements per 128-byte swizzle row
  static constexpr int NPART = two_part(MODE) ? 2 : 1;                   // hi + lo (tf32x3, f16x3) / hi + [hi|lo as bf16] (tf32x2)
  static constexpr int TILE_BYTES = TC_BM * 128;                         // 128 x BK (K-major) == BK x 128 (MN-major)
  static constexpr int BOX_MN_BYTES = BK * 128;                          // one MN-major box: BK rows x 128 bytes
  static constexpr int STAGE_BYTES = 2 * NPART * TILE_BYTES;
  static constexpr int STAGES = two_part(MODE) ? 3 : 6;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int NBAR = 2 * STAGES + 4;
  static constexpr int FMT = is32(MODE) ? 2 : (MODE == FBN_PREC_F16X3 ? 0 : 1);      // UMMA format: F16 = 0, BF16 = 1, TF32 = 2
  static constexpr bool SCALED = MODE == FBN_PREC_F16X3;
  // k-blocks per TMEM chunk.  tf32x2 issues 8 MMAs per k-block instead of 12: 6 k-blocks keep the 48 accumulations per chunk (same
  // truncation error) and give the epilogue warps as long to fold a chunk as before (MLP-1 forward GEMM, B = 65536: 460 us with 4,
  // 441 us with 6; tf32x3: 517 us).  The gain stays well below the 1/3 fewer MMAs because operand delivery, not the tensor pipe, then
  // sets the pace: 64 KB of smem fill per CTA and k-block in 1.2 us is ~7.7 TB/s of L2 -> smem traffic over the 148 SMs.
  static constexpr int CHUNK = MODE == FBN_PREC_TF32X2 ? 6 : TC_CHUNK;
};

// the MMAs of one UK-wide k-step (k = 0 .. BK/UK-1) of a staged k-block.  a_p0/b_p0 = part 0 (hi), a_p1/b_p1 = part 1.
template <int MODE>
__device__ __forceinline__ void issue_kstep(uint32_t tacc, uint64_t a_p0, uint64_t b_p0, uint64_t a_p1, uint64_t b_p1, uint64_t adv_a,
                                            uint64_t adv_b, int k, uint32_t idesc, uint32_t idesc_bf16, uint32_t acc) {
  if (MODE == FBN_PREC_TF32X3 || MODE == FBN_PREC_F16X3) {
    umma<MODE>(tacc, a_p1 + k * adv_a, b_p0 + k * adv_b, idesc, acc);      // small terms first
    umma<MODE>(tacc, a_p0 + k * adv_a, b_p1 + k * adv_b, idesc, 1u);
    umma<MODE>(tacc, a_p0 + k * adv_a, b_p0 + k * adv_b, idesc, 1u);
  } else if (MODE == FBN_PREC_TF32X2) {
    // K-major only: a k-step covers 8 tf32 = 32 bytes of the hi row.  The bf16 corrections run at 16 elements per MMA, i.e. once per
    // two tf32 k-steps: part-1 row = [hi_bf16 x32 (bytes 0..63) | lo_bf16 x32 (bytes 64..127)], 16 bf16 = 32 bytes = 2 x (16-byte units).
    if ((k & 1) == 0) {
      const uint64_t h = (uint64_t)(k >> 1) * 2, l = 4 + (uint64_t)(k >> 1) * 2;
      umma<FBN_PREC_BF16>(tacc, a_p1 + l, b_p1 + h, idesc_bf16, acc);                   // Al * Bh
      umma<FBN_PREC_BF16>(tacc, a_p1 + h, b_p1 + l, idesc_bf16, 1u);                    // Ah * Bl
      umma<FBN_PREC_TF32X3>(tacc, a_p0 + k * adv_a, b_p0 + k * adv_b, idesc, 1u);       // Ah * Bh (tf32)
    } else {
      umma<FBN_PREC_TF32X3>(tacc, a_p0 + k * adv_a, b_p0 + k * adv_b, idesc, acc);
    }
  } else {
    umma<MODE>(tacc, a_p0 + k * adv_a, b_p0 + k * adv_b, idesc, acc);
  }
}

struct TcMaps { CUtensorMap a[2], b[2]; };   // [0] = hi (or the only part), [1] = lo

template <int MODE, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcMaps tm, const TcArgs g) {
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
  const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM, sp = blockIdx.z % g.splits, bi = blockIdx.z / g.splits;
  if (g.nmask != ~0ull && !((g.nmask >> (n0 / 128)) & 1ull)) return;
  const int acol = bi * g.a_bc, arow = bi * g.a_br, bcol = bi * g.b_bc, brow = bi * g.b_br;

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
#pragma unroll
      for (int p = 0; p < Cfg::NPART; ++p) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.a[p])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.b[p])) : "memory");
      }
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!active(kb)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, Cfg::STAGE_BYTES);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        const int kc = kb * Cfg::BK;
#pragma unroll
        for (int p = 0; p < Cfg::NPART; ++p) {
          uint8_t* sa = st + p * Cfg::TILE_BYTES;                       // stage = [A parts | B parts]
          uint8_t* sb = st + (Cfg::NPART + p) * Cfg::TILE_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tm.a[p], full + s, kc + acol, m0 + arow);  // box: BK elements of K x 128 rows of M
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / Cfg::EPB; ++j)                   // boxes: 128 B of M x BK rows of K
              tma_load_2d(sa + j * Cfg::BOX_MN_BYTES, &tm.a[p], full + s, m0 + j * Cfg::EPB + acol, kc + arow);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tm.b[p], full + s, kc + bcol, n0 + brow);
          } else {
#pragma unroll
            for (int j = 0; j < TC_BN / Cfg::EPB; ++j)
              tma_load_2d(sb + j * Cfg::BOX_MN_BYTES, &tm.b[p], full + s, n0 + j * Cfg::EPB + bcol, kc + brow);
          }
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nact > 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::FMT, TC_BM, TC_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // per-MMA K advance of the descriptor start address (16-byte units): K-major: UK elements inside the 128-byte
      // swizzle row; MN-major: UK rows of 128 bytes
      constexpr uint64_t adv_a = A_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint64_t adv_b = B_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint32_t lbo_a = A_MN ? Cfg::BOX_MN_BYTES : 16, lbo_b = B_MN ? Cfg::BOX_MN_BYTES : 16;
      constexpr bool base32 = is32(MODE);   // 4-byte elements, MN-major: 32B-atom swizzle
      constexpr uint32_t sbo_a = (A_MN && base32) ? 512 : 1024, sbo_b = (B_MN && base32) ? 512 : 1024;
      constexpr uint32_t lay_a = (A_MN && base32) ? 1 : 2, lay_b = (B_MN && base32) ? 1 : 2;
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!active(kb)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        const int chunk = it / Cfg::CHUNK, buf = chunk & 1, pos = it % Cfg::CHUNK;
        if (pos == 0) {  // the epilogue must have drained this accumulator (two chunks ago)
          mbar_wait(tempty + buf, ((chunk >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * TC_BN);
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t a_hi = make_desc(sa, lbo_a, sbo_a, lay_a);
        const uint64_t b_hi = make_desc(sa + Cfg::NPART * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
#pragma unroll
        for (int k = 0; k < Cfg::BK / Cfg::UK; ++k) {
          const uint32_t acc = (pos > 0 || k > 0) ? 1u : 0u;
          const uint64_t a_lo = make_desc(sa + Cfg::TILE_BYTES, lbo_a, sbo_a, lay_a);           // part 1 (unused for bf16)
          const uint64_t b_lo = make_desc(sa + 3 * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
          issue_kstep<MODE>(tacc, a_hi, b_hi, a_lo, b_lo, adv_a, adv_b, k, idesc, make_idesc(1, TC_BM, TC_BN, 0, 0), acc);
        }
        tc_commit(empty + s);                                  // frees the smem stage once the MMAs have read it
        if (pos == Cfg::CHUNK - 1 || it == nact - 1) tc_commit(tfull + buf);   // chunk complete
        ++it;
      }
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 == output rows m0 + 32*(w%4) + lane
    const int q = warp & 3;
    float s0 = 1.f, s1 = 1.f;
    if (Cfg::SCALED) {
      if (g.inv_sa) s0 = __ldg(g.inv_sa);
      if (g.inv_sb) s1 = __ldg(g.inv_sb);
    }
    float acc[TC_BN];
#pragma unroll
    for (int j = 0; j < TC_BN; ++j) acc[j] = 0.f;
    const int nchunks = (nact + Cfg::CHUNK - 1) / Cfg::CHUNK;
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
    {
      const long long row0 = (long long)m0 + q * 32;
      float* cbase = g.C + (long long)bi * g.strideC + (long long)sp * g.strideSplit + row0 * g.ldc + n0;
      epilogue_store(acc, smem_u32(smem) + (uint32_t)((warp - 2) * EPI_WARP_BYTES), lane, row0, g.M, cbase, g.ldc,
                     g.bias ? g.bias + n0 : nullptr, g.accumulate, s0, s1);
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
// Persistent single-CTA kernel: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the linearised
// (batch x split x M-tile x N-tile) space.  Barriers, TMEM and the TMA/MMA rings are set up once per CTA; the producer and
// the MMA issuer run ahead across tile boundaries while the epilogue warps drain / store the previous tile (the two TMEM
// accumulators alternate per chunk exactly as inside a tile).  This is what the many small contractions of the step need
// (K = 128 bilinear transforms: 4 k-blocks per tile -- set-up and epilogue used to dominate them: 7.6 % tensor activity).
// The epilogue transposes through DEDICATED staging smem (the stages may already be refilled for the next tile).
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct TcpCfg {
  static constexpr int STAGES = two_part(MODE) ? 2 : 4;
  static constexpr int STAGE_BYTES = TcCfg<MODE>::STAGE_BYTES;
  static constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;
  static constexpr int NBAR = 2 * STAGES + 4;
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

struct TileInfo { int n0, m0, sp, bi, kb0, kb1, nact; bool skip; };

// tile order (fastest -> slowest): N tile, batch x split, M tile -- tiles that run at the same time share their A rows
// (the four bilinear transforms of a row block read adjacent 512-byte column blocks of the same rows)
template <int MODE>
__device__ __forceinline__ TileInfo tile_info(const TcArgs& g, long long t, int nN, int nZ, int kblocks, int per) {
  using Cfg = TcCfg<MODE>;
  TileInfo ti;
  const int x = (int)(t % nN);
  const long long r = t / nN;
  const int z = (int)(r % nZ), y = (int)(r / nZ);
  ti.n0 = x * TC_BN; ti.m0 = y * TC_BM; ti.sp = z % g.splits; ti.bi = z / g.splits;
  ti.skip = g.nmask != ~0ull && !((g.nmask >> (ti.n0 / 128)) & 1ull);
  ti.kb0 = ti.sp * per; ti.kb1 = min(kblocks, ti.kb0 + per);
  int nact = 0;
  for (int kb = ti.kb0; kb < ti.kb1; ++kb)
    nact += (g.kmask == ~0ull || ((g.kmask >> ((kb * Cfg::BK) / 128)) & 1ull)) ? 1 : 0;
  ti.nact = nact;
  return ti;
}

template <int MODE, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tcp_kernel(const __grid_constant__ TcMaps tm, const TcArgs g, long long ntiles,
                                                                 int nN, int nZ) {
  using Cfg = TcCfg<MODE>;
  using PC = TcpCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + PC::STAGES * PC::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + PC::EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + PC::STAGES;
  uint64_t* tfull = bars + 2 * PC::STAGES;
  uint64_t* tempty = bars + 2 * PC::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + PC::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (int)((g.K + Cfg::BK - 1) / Cfg::BK);
  const int per = (kblocks + g.splits - 1) / g.splits;
  auto active = [&](int kb) { return g.kmask == ~0ull || ((g.kmask >> ((kb * Cfg::BK) / 128)) & 1ull); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < PC::STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
#pragma unroll
      for (int p = 0; p < Cfg::NPART; ++p) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.a[p])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.b[p])) : "memory");
      }
      int it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const TileInfo ti = tile_info<MODE>(g, t, nN, nZ, kblocks, per);
        if (ti.skip) continue;
        const int acol = ti.bi * g.a_bc, arow = ti.bi * g.a_br, bcol = ti.bi * g.b_bc, brow = ti.bi * g.b_br;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          if (!active(kb)) continue;
          const int s = it % PC::STAGES;
          const uint32_t ph = (it / PC::STAGES) & 1;
          mbar_wait(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, PC::STAGE_BYTES);
          uint8_t* st = smem + s * PC::STAGE_BYTES;
          const int kc = kb * Cfg::BK;
#pragma unroll
          for (int p = 0; p < Cfg::NPART; ++p) {
            uint8_t* sa = st + p * Cfg::TILE_BYTES;
            uint8_t* sb = st + (Cfg::NPART + p) * Cfg::TILE_BYTES;
            if (!A_MN) {
              tma_load_2d(sa, &tm.a[p], full + s, kc + acol, ti.m0 + arow);
            } else {
#pragma unroll
              for (int j = 0; j < TC_BM / Cfg::EPB; ++j)
                tma_load_2d(sa + j * Cfg::BOX_MN_BYTES, &tm.a[p], full + s, ti.m0 + j * Cfg::EPB + acol, kc + arow);
            }
            if (!B_MN) {
              tma_load_2d(sb, &tm.b[p], full + s, kc + bcol, ti.n0 + brow);
            } else {
#pragma unroll
              for (int j = 0; j < TC_BN / Cfg::EPB; ++j)
                tma_load_2d(sb + j * Cfg::BOX_MN_BYTES, &tm.b[p], full + s, ti.n0 + j * Cfg::EPB + bcol, kc + brow);
            }
          }
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::FMT, TC_BM, TC_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint64_t adv_a = A_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint64_t adv_b = B_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint32_t lbo_a = A_MN ? Cfg::BOX_MN_BYTES : 16, lbo_b = B_MN ? Cfg::BOX_MN_BYTES : 16;
      constexpr bool base32 = is32(MODE);
      constexpr uint32_t sbo_a = (A_MN && base32) ? 512 : 1024, sbo_b = (B_MN && base32) ? 512 : 1024;
      constexpr uint32_t lay_a = (A_MN && base32) ? 1 : 2, lay_b = (B_MN && base32) ? 1 : 2;
      int it = 0, chunk = 0;       // global ring positions (continue across tiles)
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const TileInfo ti = tile_info<MODE>(g, t, nN, nZ, kblocks, per);
        if (ti.skip || ti.nact == 0) continue;
        int done = 0;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          if (!active(kb)) continue;
          const int s = it % PC::STAGES;
          const uint32_t ph = (it / PC::STAGES) & 1;
          const int buf = chunk & 1, pos = done % Cfg::CHUNK;
          if (pos == 0) {
            mbar_wait(tempty + buf, ((chunk >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(buf * TC_BN);
          const uint32_t sa = smem_u32(smem + s * PC::STAGE_BYTES);
          const uint64_t a_hi = make_desc(sa, lbo_a, sbo_a, lay_a);
          const uint64_t b_hi = make_desc(sa + Cfg::NPART * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
#pragma unroll
          for (int k = 0; k < Cfg::BK / Cfg::UK; ++k) {
            const uint32_t acc = (pos > 0 || k > 0) ? 1u : 0u;
            const uint64_t a_lo = make_desc(sa + Cfg::TILE_BYTES, lbo_a, sbo_a, lay_a);
            const uint64_t b_lo = make_desc(sa + 3 * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
            issue_kstep<MODE>(tacc, a_hi, b_hi, a_lo, b_lo, adv_a, adv_b, k, idesc, make_idesc(1, TC_BM, TC_BN, 0, 0), acc);
          }
          tc_commit(empty + s);
          ++done;
          ++it;
          if (pos == Cfg::CHUNK - 1 || done == ti.nact) {
            tc_commit(tfull + buf);
            ++chunk;
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const uint32_t stage = smem_u32(epi) + (uint32_t)((warp - 2) * EPI_WARP_BYTES);
    float s0 = 1.f, s1 = 1.f;
    if (Cfg::SCALED) {
      if (g.inv_sa) s0 = __ldg(g.inv_sa);
      if (g.inv_sb) s1 = __ldg(g.inv_sb);
    }
    int chunk = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const TileInfo ti = tile_info<MODE>(g, t, nN, nZ, kblocks, per);
      if (ti.skip) continue;
      float acc[TC_BN];
#pragma unroll
      for (int j = 0; j < TC_BN; ++j) acc[j] = 0.f;
      const int nchunks = (ti.nact + Cfg::CHUNK - 1) / Cfg::CHUNK;
      for (int c = 0; c < nchunks; ++c, ++chunk) {
        const int buf = chunk & 1;
        mbar_wait(tfull + buf, (chunk >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_BN + c0), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        mbar_arrive(tempty + buf);
      }
      const long long row0 = (long long)ti.m0 + q * 32;
      float* cbase = g.C + (long long)ti.bi * g.strideC + (long long)ti.sp * g.strideSplit + row0 * g.ldc + ti.n0;
      epilogue_store(acc, stage, lane, row0, g.M, cbase, g.ldc, g.bias ? g.bias + ti.n0 : nullptr, g.accumulate, s0, s1);
      __syncwarp();        // the staging rows are reused by the next tile
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
// CTA-pair kernel (cta_group::2): one 256 x 256 output tile per 2-CTA cluster.
//   Each CTA stages ITS 128 rows of A and ITS 128-row half of the B tile (64 KB per stage for tf32x3, as above) but the
//   pair's tensor cores multiply 256 x 256 per instruction, so the L2 -> smem traffic per MMA is halved -- the 1-CTA
//   kernel is bound by exactly that traffic (profiles/: ~50 % tensor-pipe activity).
//   Roles per CTA (320 threads): warp 0 TMA producer (TMA completes on the LEADER's full barrier), warp 1 of the leader
//   CTA issues tcgen05.mma.cta_group::2 and multicasts tcgen05.commit to both CTAs' barriers, warps 2..9 epilogue
//   (lane quarter = warp % 4, 128-column half = (warp - 2) / 4) with the same chunked round-to-nearest folding.
// ------------------------------------------------------------------------------------------------
constexpr int TC2_THREADS = 320;
constexpr int TC2_BN = 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // arrives on the same barrier in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <int MODE>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (MODE == FBN_PREC_TF32X3) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}

template <int MODE, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ TcMaps tm, const TcArgs g) {
  using Cfg = TcCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                            // used on the leader: TMA bytes of BOTH CTAs
  uint64_t* empty = bars + Cfg::STAGES;             // per CTA: stage free (multicast commit)
  uint64_t* tfull = bars + 2 * Cfg::STAGES;         // per CTA [2]: accumulator chunk complete (multicast commit)
  uint64_t* tempty = bars + 2 * Cfg::STAGES + 2;    // leader [2]: both CTAs' epilogues drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n0 = blockIdx.y * TC2_BN, sp = blockIdx.z;
  const int m0 = (blockIdx.x >> 1) * 256 + (int)rank * 128;      // this CTA's 128 rows of the 256-row cluster tile
  // both 128-column halves of this N tile inactive -> nothing to do (uniform across the pair)
  if (g.nmask != ~0ull && !((g.nmask >> (n0 / 128)) & 3ull)) return;

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
      mbar_init(tempty + b, 2 * 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // both CTAs, same warp id: two accumulators of 256 fp32 columns = the whole TMEM
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nact > 0) {
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!active(kb)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);                                   // my smem stage is free
        const uint32_t lbar = mapa_u32(smem_u32(full + s), 0);          // leader's full barrier
        if (rank == 0) mbar_expect_tx(full + s, 2 * Cfg::STAGE_BYTES);  // bytes of both CTAs land here
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        const int kc = kb * Cfg::BK;
        const int nb = n0 + (int)rank * 128;                            // my half of the B tile
#pragma unroll
        for (int p = 0; p < Cfg::NPART; ++p) {
          uint8_t* sa = st + p * Cfg::TILE_BYTES;
          uint8_t* sb = st + (Cfg::NPART + p) * Cfg::TILE_BYTES;
          if (!A_MN) {
            tma_load_2d_pair(sa, &tm.a[p], lbar, kc, m0);
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / Cfg::EPB; ++j) tma_load_2d_pair(sa + j * Cfg::BOX_MN_BYTES, &tm.a[p], lbar, m0 + j * Cfg::EPB, kc);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, &tm.b[p], lbar, kc, nb);
          } else {
#pragma unroll
            for (int j = 0; j < 128 / Cfg::EPB; ++j) tma_load_2d_pair(sb + j * Cfg::BOX_MN_BYTES, &tm.b[p], lbar, nb + j * Cfg::EPB, kc);
          }
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0 && nact > 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::FMT, 256, TC2_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint64_t adv_a = A_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint64_t adv_b = B_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint32_t lbo_a = A_MN ? Cfg::BOX_MN_BYTES : 16, lbo_b = B_MN ? Cfg::BOX_MN_BYTES : 16;
      constexpr bool base32 = is32(MODE);
      constexpr uint32_t sbo_a = (A_MN && base32) ? 512 : 1024, sbo_b = (B_MN && base32) ? 512 : 1024;
      constexpr uint32_t lay_a = (A_MN && base32) ? 1 : 2, lay_b = (B_MN && base32) ? 1 : 2;
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!active(kb)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        const int chunk = it / Cfg::CHUNK, buf = chunk & 1, pos = it % Cfg::CHUNK;
        if (pos == 0) {
          mbar_wait(tempty + buf, ((chunk >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * TC2_BN);
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t a_hi = make_desc(sa, lbo_a, sbo_a, lay_a);
        const uint64_t b_hi = make_desc(sa + Cfg::NPART * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
#pragma unroll
        for (int k = 0; k < Cfg::BK / Cfg::UK; ++k) {
          const uint32_t acc = (pos > 0 || k > 0) ? 1u : 0u;
          if (MODE == FBN_PREC_TF32X2) {         // see issue_kstep: tf32 hi*hi every k-step, the bf16 corrections every second one
            const uint64_t a_p1 = make_desc(sa + Cfg::TILE_BYTES, lbo_a, sbo_a, lay_a);
            const uint64_t b_p1 = make_desc(sa + 3 * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
            constexpr uint32_t idesc_bf16 = make_idesc(1, 256, TC2_BN, 0, 0);
            if ((k & 1) == 0) {
              const uint64_t h = (uint64_t)(k >> 1) * 2, l = 4 + (uint64_t)(k >> 1) * 2;
              umma_pair<FBN_PREC_BF16>(tacc, a_p1 + l, b_p1 + h, idesc_bf16, acc);
              umma_pair<FBN_PREC_BF16>(tacc, a_p1 + h, b_p1 + l, idesc_bf16, 1u);
              umma_pair<FBN_PREC_TF32X3>(tacc, a_hi + k * adv_a, b_hi + k * adv_b, idesc, 1u);
            } else {
              umma_pair<FBN_PREC_TF32X3>(tacc, a_hi + k * adv_a, b_hi + k * adv_b, idesc, acc);
            }
          } else if (Cfg::NPART == 2) {
            const uint64_t a_lo = make_desc(sa + Cfg::TILE_BYTES, lbo_a, sbo_a, lay_a);
            const uint64_t b_lo = make_desc(sa + 3 * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
            umma_pair<MODE>(tacc, a_lo + k * adv_a, b_hi + k * adv_b, idesc, acc);
            umma_pair<MODE>(tacc, a_hi + k * adv_a, b_lo + k * adv_b, idesc, 1u);
            umma_pair<MODE>(tacc, a_hi + k * adv_a, b_hi + k * adv_b, idesc, 1u);
          } else {
            umma_pair<MODE>(tacc, a_hi + k * adv_a, b_hi + k * adv_b, idesc, acc);
          }
        }
        tc_commit_pair(empty + s);
        if (pos == Cfg::CHUNK - 1 || it == nact - 1) tc_commit_pair(tfull + buf);
        ++it;
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    float s0 = 1.f, s1 = 1.f;
    if (Cfg::SCALED) {
      if (g.inv_sa) s0 = __ldg(g.inv_sa);
      if (g.inv_sb) s1 = __ldg(g.inv_sb);
    }
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.f;
    const int nchunks = (nact + Cfg::CHUNK - 1) / Cfg::CHUNK;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(tfull + buf, (c >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC2_BN + half * 128 + c0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(smem_u32(tempty + buf), 0));       // tell the leader's MMA warp
    }
    const int ncol0 = n0 + half * 128;
    const bool col_on = ncol0 < g.N && (g.nmask == ~0ull || ((g.nmask >> (ncol0 / 128)) & 1ull));
    if (col_on) {
      const long long row0 = (long long)m0 + q * 32;
      float* cbase = g.C + (long long)sp * g.strideSplit + row0 * g.ldc + ncol0;
      epilogue_store(acc, smem_u32(smem) + (uint32_t)((warp - 2) * EPI_WARP_BYTES), lane, row0, g.M, cbase, g.ldc,
                     g.bias ? g.bias + ncol0 : nullptr, g.accumulate, s0, s1);
    }
  }
  tc_fence_before();
  cluster_sync_all();          // the peer's smem / TMEM must stay alive until every MMA and remote arrive has landed
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent CTA-pair kernel: 74 clusters of 2 CTAs (one CTA per SM) walk the tile list
//   t = cluster, cluster + #clusters, ...  over  (M tile of 256) x (split) x (N tile of 2 live 128-column blocks), N fastest,
// so that (i) barriers / TMEM / tensor maps are set up once per SM instead of once per tile, (ii) the TMA producer and the MMA
// issuer run ahead into tile i+1 while the epilogue warps fold and store tile i (with K = 512 the non-overlapped prologue +
// epilogue of the one-tile-per-cluster kernel cost ~1/3 of the tile: 65 % tensor activity in the MLP-1 data gradient),
// (iii) clusters that run at the same time work on the same A rows (the two N tiles of an M block are adjacent in the order, so
// the second one finds A in L2: the one-tile kernel re-read the whole packed A from DRAM once per N tile), and (iv) the N tiles
// are built from the LIVE 128-column blocks only: each CTA of a pair stages its own 128-column half of the B tile, so any two
// live blocks can share a tile -- the 15 live blocks of the MLP input (of 21) fill 8 tiles instead of the 9 (3 of them half dead)
// that a regular 256-column grid needs.
// The epilogue stages through a DEDICATED 4 KB per warp (32 rows x 32 floats, XOR-swizzled) because the pipeline stages
// are already being refilled for the next tile.
// ------------------------------------------------------------------------------------------------
struct Tc2pArgs {
  TcArgs g;
  long long ntiles;
  int nNt;                  // N tiles = ceil(nlive / 2)
  int nlive;                // live 128-column blocks
  unsigned char nb[64];     // their block indices
};

constexpr int EPI2_WARP_FLOATS = 32 * 32;

// stage: shared-space byte address of this warp's 4 KB (32 rows x 32 floats, 16-byte chunks XOR-swizzled by the row)
__device__ __forceinline__ void epilogue_store_sw(const float (&acc)[128], uint32_t stage, int lane, long long row0, long long M,
                                                  float* cbase, long long ldc, const float* bias, int accumulate, float s0 = 1.f,
                                                  float s1 = 1.f) {
  const int rr = lane >> 3, cc = lane & 7;
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4)
      sts128(stage + (uint32_t)(lane * 32 + ((c4 ^ (lane & 7)) << 2)) * 4u,
             make_float4(acc[q4 * 32 + c4 * 4], acc[q4 * 32 + c4 * 4 + 1], acc[q4 * 32 + c4 * 4 + 2], acc[q4 * 32 + c4 * 4 + 3]));
    __syncwarp();
    const float4 bv = bias ? ld4(bias + q4 * 32 + cc * 4) : f4(0.f);
    float4 o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {       // all eight staged rows first, then the eight global stores
      const int r = i * 4 + rr;
      o[i] = lds128(stage + (uint32_t)(r * 32 + ((cc ^ (r & 7)) << 2)) * 4u);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + rr;
      if (row0 + r < M) {
        float4 x = (o[i] * s0) * s1 + bv;
        float* cp = cbase + (long long)r * ldc + q4 * 32 + cc * 4;
        if (accumulate) x += *reinterpret_cast<const float4*>(cp);
        st4(cp, x);
      }
    }
    __syncwarp();
  }
}

struct Tile2 { int m0, sp, nbA, nbB, kb0, kb1, nact; bool hasB; };

template <int MODE>
__device__ __forceinline__ Tile2 tile2_info(const Tc2pArgs& p, long long t, int kblocks, int per) {
  using Cfg = TcCfg<MODE>;
  Tile2 ti;
  const int x = (int)(t % p.nNt);
  const long long r = t / p.nNt;
  ti.sp = (int)(r % p.g.splits);
  ti.m0 = (int)(r / p.g.splits) * 256;
  ti.nbA = p.nb[2 * x];
  ti.hasB = 2 * x + 1 < p.nlive;
  ti.nbB = ti.hasB ? p.nb[2 * x + 1] : p.nb[2 * x];      // dead second half: stage the first block again, never stored
  ti.kb0 = ti.sp * per; ti.kb1 = min(kblocks, ti.kb0 + per);
  int nact = 0;
  for (int kb = ti.kb0; kb < ti.kb1; ++kb)
    nact += (p.g.kmask == ~0ull || ((p.g.kmask >> ((kb * Cfg::BK) / 128)) & 1ull)) ? 1 : 0;
  ti.nact = nact;
  return ti;
}

template <int MODE>
struct Tc2pCfg {
  static constexpr int STAGES = TcCfg<MODE>::STAGES;
  static constexpr int STAGE_BYTES = TcCfg<MODE>::STAGE_BYTES;
  static constexpr int EPI_BYTES = 8 * EPI2_WARP_FLOATS * 4;        // 32 KB
  static constexpr int NBAR = 2 * STAGES + 4;
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

template <int MODE, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
    gemm_tc2p_kernel(const __grid_constant__ TcMaps tm, const __grid_constant__ Tc2pArgs p) {
  using Cfg = TcCfg<MODE>;
  using PC = Tc2pCfg<MODE>;
  const TcArgs& g = p.g;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + PC::STAGES * PC::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + PC::EPI_BYTES);
  uint64_t* full = bars;                            // leader: TMA bytes of BOTH CTAs
  uint64_t* empty = bars + PC::STAGES;              // per CTA: stage free (multicast commit)
  uint64_t* tfull = bars + 2 * PC::STAGES;          // per CTA [2]: accumulator chunk complete (multicast commit)
  uint64_t* tempty = bars + 2 * PC::STAGES + 2;     // leader [2]: both CTAs' epilogues drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + PC::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int kblocks = (int)((g.K + Cfg::BK - 1) / Cfg::BK);
  const int per = (kblocks + g.splits - 1) / g.splits;
  auto active = [&](int kb) { return g.kmask == ~0ull || ((g.kmask >> ((kb * Cfg::BK) / 128)) & 1ull); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < PC::STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, 2 * 8);       // one arrival per epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < Cfg::NPART; ++q) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.a[q])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.b[q])) : "memory");
      }
      int it = 0;
      for (long long t = cid; t < p.ntiles; t += ncl) {
        const Tile2 ti = tile2_info<MODE>(p, t, kblocks, per);
        const int m0 = ti.m0 + (int)rank * 128;                       // this CTA's 128 rows of the 256-row tile
        const int nb = (rank == 0 ? ti.nbA : ti.nbB) * 128;           // and its 128-column half of the B tile
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          if (!active(kb)) continue;
          const int s = it % PC::STAGES;
          const uint32_t ph = (it / PC::STAGES) & 1;
          mbar_wait(empty + s, ph ^ 1);
          const uint32_t lbar = mapa_u32(smem_u32(full + s), 0);
          if (rank == 0) mbar_expect_tx(full + s, 2 * PC::STAGE_BYTES);
          uint8_t* st = smem + s * PC::STAGE_BYTES;
          const int kc = kb * Cfg::BK;
#pragma unroll
          for (int q = 0; q < Cfg::NPART; ++q) {
            uint8_t* sa = st + q * Cfg::TILE_BYTES;
            uint8_t* sb = st + (Cfg::NPART + q) * Cfg::TILE_BYTES;
            if (!A_MN) {
              tma_load_2d_pair(sa, &tm.a[q], lbar, kc, m0);
            } else {
#pragma unroll
              for (int j = 0; j < TC_BM / Cfg::EPB; ++j) tma_load_2d_pair(sa + j * Cfg::BOX_MN_BYTES, &tm.a[q], lbar, m0 + j * Cfg::EPB, kc);
            }
            if (!B_MN) {
              tma_load_2d_pair(sb, &tm.b[q], lbar, kc, nb);
            } else {
#pragma unroll
              for (int j = 0; j < 128 / Cfg::EPB; ++j) tma_load_2d_pair(sb + j * Cfg::BOX_MN_BYTES, &tm.b[q], lbar, nb + j * Cfg::EPB, kc);
            }
          }
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::FMT, 256, TC2_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint64_t adv_a = A_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint64_t adv_b = B_MN ? (uint64_t)(Cfg::UK * 128) >> 4 : (uint64_t)(Cfg::UK * Cfg::ESZ) >> 4;
      constexpr uint32_t lbo_a = A_MN ? Cfg::BOX_MN_BYTES : 16, lbo_b = B_MN ? Cfg::BOX_MN_BYTES : 16;
      constexpr bool base32 = is32(MODE);
      constexpr uint32_t sbo_a = (A_MN && base32) ? 512 : 1024, sbo_b = (B_MN && base32) ? 512 : 1024;
      constexpr uint32_t lay_a = (A_MN && base32) ? 1 : 2, lay_b = (B_MN && base32) ? 1 : 2;
      int it = 0, chunk = 0;
      for (long long t = cid; t < p.ntiles; t += ncl) {
        const Tile2 ti = tile2_info<MODE>(p, t, kblocks, per);
        if (ti.nact == 0) continue;
        int done = 0;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          if (!active(kb)) continue;
          const int s = it % PC::STAGES;
          const uint32_t ph = (it / PC::STAGES) & 1;
          const int buf = chunk & 1, pos = done % Cfg::CHUNK;
          if (pos == 0) {
            mbar_wait(tempty + buf, ((chunk >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(buf * TC2_BN);
          const uint32_t sa = smem_u32(smem + s * PC::STAGE_BYTES);
          const uint64_t a_hi = make_desc(sa, lbo_a, sbo_a, lay_a);
          const uint64_t b_hi = make_desc(sa + Cfg::NPART * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
#pragma unroll
          for (int k = 0; k < Cfg::BK / Cfg::UK; ++k) {
            const uint32_t acc = (pos > 0 || k > 0) ? 1u : 0u;
            if (Cfg::NPART == 2) {
              const uint64_t a_lo = make_desc(sa + Cfg::TILE_BYTES, lbo_a, sbo_a, lay_a);
              const uint64_t b_lo = make_desc(sa + 3 * Cfg::TILE_BYTES, lbo_b, sbo_b, lay_b);
              umma_pair<MODE>(tacc, a_lo + k * adv_a, b_hi + k * adv_b, idesc, acc);      // small terms first
              umma_pair<MODE>(tacc, a_hi + k * adv_a, b_lo + k * adv_b, idesc, 1u);
              umma_pair<MODE>(tacc, a_hi + k * adv_a, b_hi + k * adv_b, idesc, 1u);
            } else {
              umma_pair<MODE>(tacc, a_hi + k * adv_a, b_hi + k * adv_b, idesc, acc);
            }
          }
          tc_commit_pair(empty + s);
          ++done;
          ++it;
          if (pos == Cfg::CHUNK - 1 || done == ti.nact) {
            tc_commit_pair(tfull + buf);
            ++chunk;
          }
        }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t stage = smem_u32(epi) + (uint32_t)((warp - 2) * EPI2_WARP_FLOATS * 4);
    float s0 = 1.f, s1 = 1.f;
    if (Cfg::SCALED) {
      if (g.inv_sa) s0 = __ldg(g.inv_sa);
      if (g.inv_sb) s1 = __ldg(g.inv_sb);
    }
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(tempty), 0);
    int chunk = 0;
    for (long long t = cid; t < p.ntiles; t += ncl) {
      const Tile2 ti = tile2_info<MODE>(p, t, kblocks, per);
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      const int nchunks = (ti.nact + Cfg::CHUNK - 1) / Cfg::CHUNK;
      for (int c = 0; c < nchunks; ++c, ++chunk) {
        const int buf = chunk & 1;
        mbar_wait(tfull + buf, (chunk >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC2_BN + half * 128 + c0), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
        }
        // tcgen05.wait::ld is warp-collective: once it has returned the whole warp's slice of the accumulator is in registers,
        // so ONE remote arrival per warp releases the buffer (16 per chunk on the leader's barrier instead of 512)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader0 + (uint32_t)(buf * 8));
      }
      if (half == 0 || ti.hasB) {
        const int ncol0 = (half == 0 ? ti.nbA : ti.nbB) * 128;
        if (ncol0 < g.N) {
          const long long row0 = (long long)ti.m0 + (long long)rank * 128 + q * 32;
          float* cbase = g.C + (long long)ti.sp * g.strideSplit + row0 * g.ldc + ncol0;
          epilogue_store_sw(acc, stage, lane, row0, g.M, cbase, g.ldc, g.bias ? g.bias + ncol0 : nullptr, g.accumulate, s0, s1);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// pack kernel: fp32 (rows x cols, ld) -> same layout in operand format, pitch Kp elements
//   tf32x3: hi at dst, lo at dst + lo_off floats ; bf16: dst (rows x pitch) bf16
//   colmask: 128-column blocks to convert (structurally-zero blocks of the MLP input are never read)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <int MODE>
__global__ void pack_rows_kernel(const float* __restrict__ src, long long ld, long long R, long long K, long long Kp, void* dst,
                                 long long lo_off, unsigned long long colmask) {
  const long long q = Kp / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * q; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / q, c = (i % q) * 4;
    if (colmask != ~0ull && !((colmask >> (c / 128)) & 1ull)) continue;
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
      st4(d + lo_off + r * Kp + c, v - h);
    } else if (MODE == FBN_PREC_TF32X2) {
      // part 0: hi (tf32 in an fp32 word); part 1: per 32-element k-block one 128-byte row [hi as bf16 x32 | lo as bf16 x32]
      float* d = static_cast<float*>(dst);
      const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      st4(d + r * Kp + c, h);
      const float4 l = v - h;
      char* blk = reinterpret_cast<char*>(d + lo_off + r * Kp + (c & ~31LL)) + (c & 31) * 2;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(h.x, h.y), h1 = __floats2bfloat162_rn(h.z, h.w);
      __nv_bfloat162 l0 = __floats2bfloat162_rn(l.x, l.y), l1 = __floats2bfloat162_rn(l.z, l.w);
      uint2 oh, ol;
      oh.x = *reinterpret_cast<uint32_t*>(&h0); oh.y = *reinterpret_cast<uint32_t*>(&h1);
      ol.x = *reinterpret_cast<uint32_t*>(&l0); ol.y = *reinterpret_cast<uint32_t*>(&l1);
      *reinterpret_cast<uint2*>(blk) = oh;
      *reinterpret_cast<uint2*>(blk + 64) = ol;
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

// ------------------------------------------------------------------------------------------------
// FBN_PREC_F16X3 operands: one power-of-two scale per tensor.
//   pass 1 (amax_partial_kernel): per-block maxima of |x| over the converted columns -> tail[F16_REC_FLOATS + block].  Skipped when
//                                 the kernel that produced the tensor already published them (PackDst::tail, `producer_amax`).
//   pass 2 (pack_f16x3_kernel)  : every block folds the partial maxima (max is order-independent: deterministic, no atomics, nothing
//                                 to re-arm), derives s = 2^(14 - floor(log2 amax)) and writes hi = fp16_rn(s x) at dst, lo =
//                                 fp16_rn(s x - hi) at dst + lo_off; block 0 records {s, 1/s, amax} in tail[0..2] for the GEMM epilogues.
// (layout of the record and the device helpers: common.cuh)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) amax_partial_kernel(const float* __restrict__ src, long long ld, long long R, long long K,
                                                           long long Kp, unsigned long long colmask, float* __restrict__ tail) {
  const long long q = Kp / 4;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * q; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / q, c = (i % q) * 4;
    if (colmask != ~0ull && !((colmask >> (c / 128)) & 1ull)) continue;
    if (c + 3 < K) m = fmaxf(m, amax4(ld4(src + r * ld + c)));
    else
      for (long long j = c; j < K; ++j) m = fmaxf(m, fabsf(src[r * ld + j]));
  }
  // NaN inputs: fmaxf drops them here, the packed values (and every product) still carry them
  f16x3_publish_amax(m, tail);
}

__global__ void __launch_bounds__(256) pack_f16x3_kernel(const float* __restrict__ src, long long ld, long long R, long long K,
                                                         long long Kp, __half* __restrict__ dst, long long lo_off,
                                                         unsigned long long colmask, float* __restrict__ tail, int npartial) {
  const float sc = f16x3_block_scale(tail, npartial);
  const long long q = Kp / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * q; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / q, c = (i % q) * 4;
    if (colmask != ~0ull && !((colmask >> (c / 128)) & 1ull)) continue;
    float4 v = f4(0.f);
    if (c + 3 < K) v = ld4s(src + r * ld + c);
    else {
      if (c < K) v.x = src[r * ld + c];
      if (c + 1 < K) v.y = src[r * ld + c + 1];
      if (c + 2 < K) v.z = src[r * ld + c + 2];
    }
    store_f16x3_4(dst, lo_off, r * Kp + c, v, sc);
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

// 2-D map over a row-major (rows, cols) operand part with `pitch` elements per row, 128B swizzle, OOB -> 0.
//   k_major: rows index M/N, cols index K -> box = (128 bytes of K) x 128 rows
//   mn_major: rows index K, cols index M/N -> box = (128 bytes of M/N) x BK rows
static int make_map(CUtensorMap* m, int mode, const void* base, long long rows, long long cols, long long pitch, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  FBN_REQUIRE(enc != nullptr, FBN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int esz = is32(mode) ? 4 : 2;
  const cuuint32_t epb = 128 / esz;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(pitch * esz)};
  cuuint32_t box[2] = {epb, mn_major ? epb : 128u};   // MN-major: BK == epb rows of K
  cuuint32_t estr[2] = {1, 1};
  FBN_REQUIRE(aligned16(base) && (pitch * esz) % 16 == 0, FBN_ERR_ALIGN, "tcgen05 operand is not 16-byte aligned");
  CUresult r = enc(m, is32(mode) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (mode == FBN_PREC_F16X3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (mn_major && esz == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FBN_REQUIRE(r == CUDA_SUCCESS, FBN_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows %lld cols %lld pitch %lld)", (int)r, rows,
              cols, pitch);
  return FBN_OK;
}

static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

size_t packed_bytes(long long rows, long long cols, int precision) {
  const long long pitch = round_up(cols, precision == FBN_PREC_TF32X2 ? 32 : 8);
  if (precision == FBN_PREC_F16X3)      // hi | lo fp16 + the scale record and the partial maxima of the amax pass
    return (size_t)(rows * pitch * 4) + 1024 + (size_t)(F16_REC_FLOATS + F16_AMAX_BLOCKS) * sizeof(float);
  return (size_t)(rows * pitch * (is32(precision) ? 8 : 2)) + 1024;
}

Packed packed_describe(void* region, long long rows, long long cols, int precision) {
  Packed p;
  p.data = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(region) + 1023) & ~uintptr_t(1023));
  p.pitch = round_up(cols, 8);
  p.lo_off = rows * p.pitch;
  p.rows = rows; p.cols = cols;
  if (precision == FBN_PREC_F16X3) p.scale = reinterpret_cast<float*>(static_cast<char*>(p.data) + (size_t)rows * p.pitch * 4);
  return p;
}

// converts src (rows x cols fp32, ld) into operand format at dst (1024-byte aligned inside the caller's region)
int pack_operand(const float* src, long long ld, long long rows, long long cols, int precision, void* dst, unsigned long long colmask,
                 Packed* out, cudaStream_t st, bool producer_amax) {
  FBN_REQUIRE(precision == FBN_PREC_TF32X3 || precision == FBN_PREC_BF16 || precision == FBN_PREC_TF32X2 || precision == FBN_PREC_F16X3,
              FBN_ERR_ARG, "pack_operand: bad precision");
  FBN_REQUIRE(aligned16(src) && ld % 4 == 0, FBN_ERR_ALIGN, "pack_operand: source must be 16-byte aligned with ld %% 4 == 0");
  *out = packed_describe(dst, rows, cols, precision);
  if (precision == FBN_PREC_F16X3) {
    const long long n4 = rows * (out->pitch / 4);
    int nb1 = -1;       // producer_amax: the count is in the record
    if (!producer_amax) {
      nb1 = (int)std::max<long long>(1, std::min<long long>(cdiv(n4, 256 * 4), std::min(F16_AMAX_BLOCKS, 4 * num_sms())));
      amax_partial_kernel<<<nb1, 256, 0, st>>>(src, ld, rows, cols, out->pitch, colmask, out->scale);
      FBN_CHECK_LAUNCH();
    }
    const int nb2 = (int)std::max<long long>(1, std::min<long long>(cdiv(n4, 256), 16LL * num_sms()));
    pack_f16x3_kernel<<<nb2, 256, 0, st>>>(src, ld, rows, cols, out->pitch, static_cast<__half*>(out->data), out->lo_off, colmask,
                                           out->scale, nb1);
    FBN_CHECK_LAUNCH();
    return FBN_OK;
  }
  if (precision == FBN_PREC_TF32X2) {          // whole 32-element k-blocks per row (the interleaved bf16 part needs them)
    out->pitch = round_up(cols, 32);
    out->lo_off = rows * out->pitch;
  }
  const long long pitch = out->pitch, lo_off = out->lo_off;
  void* base = out->data;
  const long long n = rows * (pitch / 4);
  int blocks = (int)std::min<long long>(cdiv(n, 256), 16LL * num_sms());
  if (precision == FBN_PREC_TF32X3)
    pack_rows_kernel<FBN_PREC_TF32X3><<<std::max(blocks, 1), 256, 0, st>>>(src, ld, rows, cols, pitch, base, lo_off, colmask);
  else if (precision == FBN_PREC_TF32X2)
    pack_rows_kernel<FBN_PREC_TF32X2><<<std::max(blocks, 1), 256, 0, st>>>(src, ld, rows, cols, pitch, base, lo_off, colmask);
  else
    pack_rows_kernel<FBN_PREC_BF16><<<std::max(blocks, 1), 256, 0, st>>>(src, ld, rows, cols, pitch, base, lo_off, colmask);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

// bytes of scratch one gemm_tc call needs when it has to pack both operands itself
size_t gemm_tc_scratch_bytes(long long M, long long N, long long K, int precision) {
  return packed_bytes(M, K, precision) + packed_bytes(N, K, precision) + packed_bytes(K, std::max(M, N), precision) + 4096;
}

bool gemm_tc_supported(const GemmArgs& g, int precision) {
  if (precision == FBN_PREC_TF32X2)   // K-major x K-major only, no pre-packed operands of another format
    return g.a_t == 0 && g.b_t != 0 && g.N % 128 == 0 && g.ldc % 4 == 0 && g.batch == 1;
  return (precision == FBN_PREC_TF32X3 || precision == FBN_PREC_BF16 || precision == FBN_PREC_F16X3) && g.N % 128 == 0 && g.ldc % 4 == 0 &&
         g.batch >= 1;
}

// fbn_set_option("tc_persistent", v): 1 = every single-CTA launch runs the persistent tile loop (epilogue of tile i overlapped with
// the loads / MMAs of tile i+1), -1 = never, 0 (default) = per-launch heuristic below.
// Measured on B200 (tools/shortk_probe.py, tools/stage_probe.py, B = 65536): the four batched bilinear transforms (K = N = 128)
// 115 -> 86 us (tf32x3), 75 -> 43 us (bf16); bf16 short-K GEMMs beat even the CTA-pair kernel (data gradient 1: 394 -> 279 us,
// MLP-2 forward 51 -> 41, data gradient 2: 70 -> 45) because with one bf16 pass the epilogue, not the tensor pipe, is the limit;
// tf32x3 long-K GEMMs lose (MLP-1 forward 533 -> 727 us), so they keep one tile per CTA / CTA pairs.
static int g_tc_persistent = 0;
void set_tc_persistent(int on) { g_tc_persistent = on; }
static bool use_persistent(bool heuristic) { return g_tc_persistent > 0 || (g_tc_persistent == 0 && heuristic); }

template <int MODE, bool A_MN, bool B_MN>
static int launch_tcp(const TcMaps& maps, const TcArgs& t, dim3 grid, cudaStream_t st) {
  using PC = TcpCfg<MODE>;
  static bool attr = false;
  if (!attr) {
    FBN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcp_kernel<MODE, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC::SMEM));
    attr = true;
  }
  const long long ntiles = (long long)grid.x * grid.y * grid.z;
  const int ctas = (int)std::min<long long>(ntiles, num_sms());
  gemm_tcp_kernel<MODE, A_MN, B_MN><<<ctas, TC_THREADS, PC::SMEM, st>>>(maps, t, ntiles, (int)grid.x, (int)grid.z);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

template <int MODE, bool A_MN, bool B_MN>
static int launch_tc(const TcMaps& maps, const TcArgs& t, dim3 grid, cudaStream_t st, bool persist = false) {
  using Cfg = TcCfg<MODE>;
  if (use_persistent(persist)) return launch_tcp<MODE, A_MN, B_MN>(maps, t, grid, st);
  static bool attr = false;
  if (!attr) {
    FBN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  gemm_tc_kernel<MODE, A_MN, B_MN><<<grid, TC_THREADS, Cfg::SMEM, st>>>(maps, t);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

template <int MODE, bool A_MN, bool B_MN>
static int launch_tc2(const TcMaps& maps, const TcArgs& t, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<MODE>;
  static bool attr = false;
  if (!attr) {
    FBN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<MODE, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  gemm_tc2_kernel<MODE, A_MN, B_MN><<<grid, TC2_THREADS, Cfg::SMEM, st>>>(maps, t);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

static int g_tc_reserve_sms = 0;
void set_tc_reserve_sms(int n) { g_tc_reserve_sms = std::max(0, std::min(n, 64)); }
int tc_reserved_sms() { return g_tc_reserve_sms; }

template <int MODE, bool A_MN, bool B_MN>
static int launch_tc2p(const TcMaps& maps, const TcArgs& t, cudaStream_t st) {
  using PC = Tc2pCfg<MODE>;
  static bool attr = false;
  if (!attr) {
    FBN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2p_kernel<MODE, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC::SMEM));
    attr = true;
  }
  Tc2pArgs p;
  p.g = t;
  p.nlive = 0;
  for (int i = 0; i < (int)(t.N / 128); ++i)
    if (t.nmask == ~0ull || ((t.nmask >> i) & 1ull)) p.nb[p.nlive++] = (unsigned char)i;
  if (p.nlive == 0) return FBN_OK;
  p.nNt = (p.nlive + 1) / 2;
  p.ntiles = cdiv(t.M, 256) * t.splits * p.nNt;
  // fbn_set_option("tc_reserve_sms", n): leave n SMs to a concurrently running collective (its CTAs cannot share an SM with
  // a 225 KB-smem GEMM CTA; a persistent grid that oversubscribes the free SMs would wait for the collective to finish)
  const int clusters = (int)std::min<long long>(p.ntiles, std::max(1, (num_sms() - g_tc_reserve_sms) / 2));
  gemm_tc2p_kernel<MODE, A_MN, B_MN><<<2 * clusters, TC2_THREADS, PC::SMEM, st>>>(maps, p);
  FBN_CHECK_LAUNCH();
  return FBN_OK;
}

static int g_f16_persist_k = 0;   // 0: the tf32x3 rule
void set_f16_persist_k(int k) { g_f16_persist_k = k; }
static int g_tc_pair = 1;   // 0: always use the 1-CTA kernel (fbn_set_option("tc_pair", 0))
void set_tc_pair(int on) { g_tc_pair = on; }
static int g_tc_pair_persistent = 1;   // 0: one 256 x 256 tile per cluster (fbn_set_option("tc_pair_persistent", 0)), for A/B runs
void set_tc_pair_persistent(int on) { g_tc_pair_persistent = on; }

template <int MODE>
static int gemm_tc_mode(const GemmArgs& g, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  using Cfg = TcCfg<MODE>;
  const bool a_mn = g.a_t != 0;   // A stored (K,M): contraction index leads -> MN-major
  const bool b_mn = g.b_t == 0;   // B stored (K,N)
  uint8_t* sp = static_cast<uint8_t*>(scratch);
  size_t left = scratch_bytes;
  // batched launch: pre-packed operands whose batch stride is a (row, column) offset inside one packed tensor become
  // TMA coordinate offsets, so all batch x split CTAs run in ONE grid (the four bilinear transforms / weight gradients)
  if (g.batch > 1 && g.batch <= 64 && g.pkA.data && g.pkB.data && g.N == TC_BN) {
    const long long ra = a_mn ? g.K : g.M, ca = a_mn ? g.M : g.K, rb = b_mn ? g.K : g.N, cb = b_mn ? g.N : g.K;
    const long long a_br = g.strideA / g.pkA.pitch, a_bc = g.strideA % g.pkA.pitch;
    const long long b_br = g.strideB / g.pkB.pitch, b_bc = g.strideB % g.pkB.pitch;
    const bool ok = (a_mn || a_bc == 0 || g.K % Cfg::BK == 0) && (b_mn || b_bc == 0 || g.K % Cfg::BK == 0) &&
                    a_bc + ca <= g.pkA.pitch + (a_bc ? 0 : 0) && b_bc + cb <= g.pkB.pitch;
    if (ok) {
      TcMaps maps;
      const long long nb1 = g.batch - 1;
      for (int p = 0; p < Cfg::NPART; ++p) {
        int rc = make_map(&maps.a[p], MODE, static_cast<uint8_t*>(g.pkA.data) + (size_t)p * g.pkA.lo_off * Cfg::ESZ, ra + nb1 * a_br,
                          std::min(ca + nb1 * a_bc, g.pkA.pitch), g.pkA.pitch, a_mn);
        if (rc) return rc;
        rc = make_map(&maps.b[p], MODE, static_cast<uint8_t*>(g.pkB.data) + (size_t)p * g.pkB.lo_off * Cfg::ESZ, rb + nb1 * b_br,
                      std::min(cb + nb1 * b_bc, g.pkB.pitch), g.pkB.pitch, b_mn);
        if (rc) return rc;
      }
      if (Cfg::NPART == 1) { maps.a[1] = maps.a[0]; maps.b[1] = maps.b[0]; }
      TcArgs t;
      t.C = g.C; t.bias = g.bias; t.M = g.M; t.N = g.N; t.K = g.K; t.ldc = g.ldc; t.splits = g.splits;
      t.strideSplit = g.strideSplit; t.kmask = g.kmask; t.nmask = g.nmask; t.accumulate = g.accumulate;
      t.a_bc = (int)a_bc; t.a_br = (int)a_br; t.b_bc = (int)b_bc; t.b_br = (int)b_br; t.strideC = g.strideC;
      if (Cfg::SCALED) { t.inv_sa = g.pkA.scale ? g.pkA.scale + 1 : nullptr; t.inv_sb = g.pkB.scale ? g.pkB.scale + 1 : nullptr; }
      dim3 grid((unsigned)(g.N / TC_BN), (unsigned)cdiv(g.M, TC_BM), (unsigned)(g.splits * g.batch));
      // short K, many tiles (the bilinear transforms / their data gradients): the tile loop hides the epilogue
      const bool persist = g.K <= 256 && g.splits == 1 && (long long)grid.x * grid.y * grid.z >= 2LL * num_sms();
      if (a_mn && b_mn) return launch_tc<MODE, true, true>(maps, t, grid, st, persist);
      if (a_mn) return launch_tc<MODE, true, false>(maps, t, grid, st, persist);
      if (b_mn) return launch_tc<MODE, false, true>(maps, t, grid, st, persist);
      return launch_tc<MODE, false, false>(maps, t, grid, st, persist);
    }
  }
  for (int bi = 0; bi < g.batch; ++bi) {
    Packed pa, pb;
    const long long ra = a_mn ? g.K : g.M, ca = a_mn ? g.M : g.K;
    const long long rb = b_mn ? g.K : g.N, cb = b_mn ? g.N : g.K;
    uint8_t* cur = sp;
    size_t rem = left;
    if (g.pkA.data) {
      pa = g.pkA;
      pa.data = static_cast<uint8_t*>(pa.data) + (size_t)bi * g.strideA * Cfg::ESZ;
    } else {
      const size_t need = packed_bytes(ra, ca, MODE);
      FBN_REQUIRE(cur != nullptr && rem >= need, FBN_ERR_ARG, "tcgen05 GEMM: operand scratch too small (%zu < %zu)", rem, need);
      int rc = pack_operand(g.A + bi * g.strideA, g.lda, ra, ca, MODE, cur, ~0ull, &pa, st);
      if (rc) return rc;
      cur += need; rem -= need;
    }
    if (g.pkB.data) {
      pb = g.pkB;
      pb.data = static_cast<uint8_t*>(pb.data) + (size_t)bi * g.strideB * Cfg::ESZ;
    } else {
      const size_t need = packed_bytes(rb, cb, MODE);
      FBN_REQUIRE(cur != nullptr && rem >= need, FBN_ERR_ARG, "tcgen05 GEMM: operand scratch too small (%zu < %zu)", rem, need);
      int rc = pack_operand(g.B + bi * g.strideB, g.ldb, rb, cb, MODE, cur, ~0ull, &pb, st);
      if (rc) return rc;
    }
    TcMaps maps;
    for (int p = 0; p < Cfg::NPART; ++p) {
      int rc = make_map(&maps.a[p], MODE, static_cast<uint8_t*>(pa.data) + (size_t)p * pa.lo_off * Cfg::ESZ, ra, ca, pa.pitch, a_mn);
      if (rc) return rc;
      rc = make_map(&maps.b[p], MODE, static_cast<uint8_t*>(pb.data) + (size_t)p * pb.lo_off * Cfg::ESZ, rb, cb, pb.pitch, b_mn);
      if (rc) return rc;
    }
    if (Cfg::NPART == 1) { maps.a[1] = maps.a[0]; maps.b[1] = maps.b[0]; }
    TcArgs t;
    t.C = g.C + bi * g.strideC; t.bias = g.bias; t.M = g.M; t.N = g.N; t.K = g.K; t.ldc = g.ldc; t.splits = g.splits;
    t.strideSplit = g.strideSplit; t.kmask = g.kmask; t.nmask = g.nmask; t.accumulate = g.accumulate;
    if (Cfg::SCALED) {
      FBN_REQUIRE(pa.scale && pb.scale, FBN_ERR_ARG, "tcgen05 GEMM (f16x3): a pre-packed operand carries no scale record");
      t.inv_sa = pa.scale + 1; t.inv_sb = pb.scale + 1;
    }
    int rc;
    // CTA pairs (256 x 256 tiles, half the L2 traffic per MMA) once they can fill most of the 148 SMs; small problems
    // keep the 128 x 128 single-CTA tiles (4x as many CTAs)
    const long long pair_ctas = 2 * cdiv(g.M, 256) * cdiv(g.N, TC2_BN) * g.splits;
    // one bf16 pass with K <= 512 is epilogue-bound: the persistent single-CTA tile loop beats the pair kernel there
    const long long single_tiles = cdiv(g.M, TC_BM) * (g.N / TC_BN);
    // (and a short-K, 128-wide tf32x3 GEMM -- the item_emb_d128 projection -- cannot pair at all: the tile loop hides its epilogue too)
    // f16x3 (three passes at the bf16 rate) sits between the two: fbn_set_option("f16_persist_k", K) moves its threshold (A/B runs)
    const bool persist = g.splits == 1 && single_tiles >= 2LL * num_sms() &&
                         (MODE == FBN_PREC_BF16 ? g.K <= 512
                                                : (MODE == FBN_PREC_F16X3 && g_f16_persist_k > 0 ? g.K <= g_f16_persist_k : (g.K <= 256 && g.N == TC_BN)));
    if (g_tc_pair && !use_persistent(persist) && g.M > 128 && g.N >= 256 && pair_ctas >= 120) {
      if (g_tc_pair_persistent && g.N / 128 <= 64) {
        if (a_mn && b_mn) rc = launch_tc2p<MODE, true, true>(maps, t, st);
        else if (a_mn) rc = launch_tc2p<MODE, true, false>(maps, t, st);
        else if (b_mn) rc = launch_tc2p<MODE, false, true>(maps, t, st);
        else rc = launch_tc2p<MODE, false, false>(maps, t, st);
        if (rc) return rc;
        continue;
      }
      dim3 grid2((unsigned)(2 * cdiv(g.M, 256)), (unsigned)cdiv(g.N, TC2_BN), (unsigned)g.splits);
      if (a_mn && b_mn) rc = launch_tc2<MODE, true, true>(maps, t, grid2, st);
      else if (a_mn) rc = launch_tc2<MODE, true, false>(maps, t, grid2, st);
      else if (b_mn) rc = launch_tc2<MODE, false, true>(maps, t, grid2, st);
      else rc = launch_tc2<MODE, false, false>(maps, t, grid2, st);
      if (rc) return rc;
      continue;
    }
    dim3 grid((unsigned)(g.N / TC_BN), (unsigned)cdiv(g.M, TC_BM), (unsigned)g.splits);
    if (a_mn && b_mn) rc = launch_tc<MODE, true, true>(maps, t, grid, st, persist);
    else if (a_mn) rc = launch_tc<MODE, true, false>(maps, t, grid, st, persist);
    else if (b_mn) rc = launch_tc<MODE, false, true>(maps, t, grid, st, persist);
    else rc = launch_tc<MODE, false, false>(maps, t, grid, st, persist);
    if (rc) return rc;
  }
  return FBN_OK;
}

// FBN_PREC_TF32X2: C[M,N] = A[M,K] * B[N,K]^T, both operands K-major, packed here (hi | interleaved bf16 hi/lo), single-CTA tiles
static int gemm_tc_x2(const GemmArgs& g, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  constexpr int MODE = FBN_PREC_TF32X2;
  const size_t na = g.pkA.data ? 0 : packed_bytes(g.M, g.K, MODE), nb = g.pkB.data ? 0 : packed_bytes(g.N, g.K, MODE);
  FBN_REQUIRE(na + nb == 0 || (scratch != nullptr && scratch_bytes >= na + nb), FBN_ERR_ARG,
              "tcgen05 GEMM (tf32x2): operand scratch too small (%zu < %zu)", scratch_bytes, na + nb);
  Packed pa = g.pkA, pb = g.pkB;      // pre-packed operands must be in the tf32x2 format (pack_operand with this precision)
  uint8_t* sp = static_cast<uint8_t*>(scratch);
  int rc = FBN_OK;
  if (!pa.data) rc = pack_operand(g.A, g.lda, g.M, g.K, MODE, sp, ~0ull, &pa, st);
  if (rc) return rc;
  if (!pb.data) rc = pack_operand(g.B, g.ldb, g.N, g.K, MODE, sp + na, ~0ull, &pb, st);
  if (rc) return rc;
  TcMaps maps;
  // part 0 clips at K (TMA zero fill); part 1 is addressed in whole 32-slot blocks, its padding was zeroed by the pack kernel
  rc = make_map(&maps.a[0], MODE, pa.data, g.M, g.K, pa.pitch, false);
  if (rc) return rc;
  rc = make_map(&maps.a[1], MODE, static_cast<float*>(pa.data) + pa.lo_off, g.M, pa.pitch, pa.pitch, false);
  if (rc) return rc;
  rc = make_map(&maps.b[0], MODE, pb.data, g.N, g.K, pb.pitch, false);
  if (rc) return rc;
  rc = make_map(&maps.b[1], MODE, static_cast<float*>(pb.data) + pb.lo_off, g.N, pb.pitch, pb.pitch, false);
  if (rc) return rc;
  TcArgs t;
  t.C = g.C; t.bias = g.bias; t.M = g.M; t.N = g.N; t.K = g.K; t.ldc = g.ldc; t.splits = g.splits;
  t.strideSplit = g.strideSplit; t.kmask = g.kmask; t.nmask = g.nmask; t.accumulate = g.accumulate;
  const long long pair_ctas = 2 * cdiv(g.M, 256) * cdiv(g.N, TC2_BN) * g.splits;
  if (g_tc_pair && g_tc_persistent <= 0 && g.M > 128 && g.N >= 256 && pair_ctas >= 120) {      // same rule as the other precisions
    dim3 grid2((unsigned)(2 * cdiv(g.M, 256)), (unsigned)cdiv(g.N, TC2_BN), (unsigned)g.splits);
    return launch_tc2<MODE, false, false>(maps, t, grid2, st);
  }
  dim3 grid((unsigned)(g.N / TC_BN), (unsigned)cdiv(g.M, TC_BM), (unsigned)g.splits);
  return launch_tc<MODE, false, false>(maps, t, grid, st, false);
}

int gemm_tc(const GemmArgs& g, int precision, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  FBN_REQUIRE(gemm_tc_supported(g, precision), FBN_ERR_SHAPE,
              "tcgen05 GEMM: unsupported shape / layout / precision (N %lld must be a multiple of 128; tf32x2 takes a_t = 0, b_t = 1 only)", g.N);
  if (g.M <= 0) return FBN_OK;
  if (precision == FBN_PREC_TF32X2) return gemm_tc_x2(g, scratch, scratch_bytes, st);
  if (precision == FBN_PREC_TF32X3) return gemm_tc_mode<FBN_PREC_TF32X3>(g, scratch, scratch_bytes, st);
  if (precision == FBN_PREC_F16X3) return gemm_tc_mode<FBN_PREC_F16X3>(g, scratch, scratch_bytes, st);
  return gemm_tc_mode<FBN_PREC_BF16>(g, scratch, scratch_bytes, st);
}

}  // namespace fbn

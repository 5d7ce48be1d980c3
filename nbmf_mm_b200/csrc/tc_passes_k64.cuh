// Tensor-core (tcgen05 / TMEM) versions of the two pass kernels: fp32, bit-packed V, 32 < K <= 64 (the library
// instantiates KB = 64 only; K <= 32 runs the two-pipeline kernels of tc_passes_k32.cuh, which are 5 % faster there:
// see the measurements in DESIGN.md section 4.1).
//
// Same math as passes.cuh, restructured like attention on Blackwell (S = QK^T -> P -> O = PV):
//   MMA1  Theta tile = resident factor tile [TMEM] . streamed factor block [smem]   (tcgen05.mma, TS form)
//   SIMT  tcgen05.ld Theta -> masked ratio in registers -> tcgen05.st back to TMEM
//   MMA2  accumulator (+)= ratio [TMEM] . streamed factor block [smem]               (tcgen05.mma, TS form)
// Both MMAs take their A operand from TMEM: an SS-form MMA with a 128-row A tile re-reads 4 KB of shared
// memory per K step and is bound by the 128 B/clk shared-memory port (40 clk per N=32 MMA, measured),
// the TS form runs at the tensor pipe's 16 clk (tools/tc_bench.cu).
//
// fp32 accuracy from tensor-core passes: x = hi + lo with hi = x truncated to tf32 (what the tensor core reads) and
// a.b ~ hi.hi + (hi.lo + lo.hi).  The first term is a kind::tf32 chain, the two correction terms (2^-11 of the sum, 8
// bits are enough) are ONE kind::f16 (bf16) chain with twice the K extent.  The streamed operand blocks are pre-split
// and pre-swizzled in global memory by format_factors.cu, so a pipeline stage is ONE 1-D bulk copy (cp.async.bulk +
// mbarrier complete_tx): no tensor maps.
//
// Structure.  The unit of the pipeline is a HALF-block: 16 rows of W in the H pass, 32 columns of H in the W pass.
// Half-block hb is dealt to SIMT pipeline hb % P; a pipeline is four warps (one per TMEM lane quarter) with its own
// Theta and ratio regions in TMEM and its own barriers, so the pipelines of a CTA sit in different phases of ld ->
// arithmetic -> st.  K <= 64 has three pipelines and ONE accumulator set fed by one MMA2 issuer in half-block order:
// that is what fits into 512 TMEM columns beside the 128-column resident operand (A 128 | Theta 3 x 16 | R 3 x 64 |
// Q, S 2 x 64 = 496).  The template also instantiates for KB = 16 / 32 (four pipelines, two accumulator sets, same
// accumulation order as tc_passes_k32.cuh: bit-identical H'), which is how the structure was A/B-tested against the
// two-pipeline kernels: 5 % slower there (more barrier traffic per entry, N = 16 MMAs at 10.3 clk instead of 8), so
// the library uses it for 32 < K <= 64 only.  The tensor core's fp32 accumulate truncates (an unbroken
// chain of 3e4 accumulations drifts by ~1e-4, measured), so each TMEM chain is kFlush blocks long and is then added
// into fp32 accumulators in shared memory / registers by the set's SIMT warps; the issuer starts the next chain (which
// overwrites the accumulators) only after they have all read the previous one (bar_flushed).
//
// Template KB = 16, 32, 64 (K extent the MMAs cover): KB = 16 uses the K <= 32 operand formats and skips the
// MMA K steps that are all padding and half of MMA2's N.
#pragma once
#include "args.h"
#include "common.cuh"
#include "tc_args.h"
#include "tc_common.cuh"

#include <type_traits>

// pointer `ptr` of the kernel arguments `a` for fit blockIdx.z of a batch (workspaces batch_stride bytes apart)
#define NBMF_TSH(ptr) batch_shift(ptr, (size_t)blockIdx.z * (size_t)a.batch_stride)

namespace nbmf {
namespace k64 {

template <int KB>
struct TcCfg {
  static_assert(KB == 16 || KB == 32 || KB == 64, "KB");
  static constexpr int KT = KB == 64 ? 64 : 32;        // K extent of the formatted operands = padded K of W, H, C, D, G
  static constexpr int P = KB == 64 ? 3 : 4;           // SIMT pipelines (4 warps each)
  static constexpr int NSETS = KB == 64 ? 1 : 2;       // accumulator sets = MMA2 issuer warps
  static constexpr int SIMT_WARPS = 4 * P;
  // MMA1 issuer warps.  Two (block b -> issuer b & 1) when each issuer then owns its pipelines and its stages
  // exclusively (P = 4: blocks of issuer e go to pipelines 2e, 2e + 1); ONE for P = 3, where consecutive uses of a
  // pipeline belong to blocks of different parity: mbarrier waits are by phase PARITY, so a waiter must never be two
  // phases away from its barrier -- a single in-order issuer requests the phases of every barrier in order, two
  // independent issuers would not (one could ask for phase u + 1 of a pipeline before phase u completed).
  static constexpr int NM1 = KB == 64 ? 1 : 2;
  static constexpr int MMA1_WARP = SIMT_WARPS;         // + issuer
  static constexpr int MMA2_WARP = SIMT_WARPS + NM1;   // + set
  static constexpr int WARPS = SIMT_WARPS + NM1 + NSETS;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int SLABS = KT / 32;                // 128-byte-wide K slabs of a K-major operand
  static constexpr int NACC = KB;                      // N of MMA2 = accumulator columns per plane
  static constexpr int KS = KB / 8;                    // tf32 K steps (8 per step) = bf16 K steps (16 bf16) of MMA1
};
constexpr int kFlush = 8;                              // blocks per TMEM accumulation chain

// ------------------------------------------------------------------------------------------------------------------
// Per-entry ratio arithmetic, two entries at a time, in PTX: the data bit of an entry is ONE predicate that drives the
// choice of x and the masking of the outputs (ptxas otherwise rebuilds a mask per entry with shifts, and the ALU pipe is
// the busiest SIMT pipe here); the adds of a pair are packed (add / sub .f32x2: one issue slot for two entries).
//   H pass:  x = (p ? theta : 1 - theta) + eps   (1 - theta saturated to [0, 1]: a Theta that rounding pushed past 1 must
//            not turn x negative), r = 1/x, hi = tf32(r), c = bf16x2(hi, r - hi) in one 32-bit word (low half = hi: the
//            K order of the bf16 correction MMA).  Planes: R = (hi, c) unmasked and Rq = the same masked by the bit of
//            `qbits` (the plane that is accumulated directly: the ones, or, per column, the zeros -- see the kernel).
// ------------------------------------------------------------------------------------------------------------------
#define NBMF_H_PAIR_BODY                                                                                   \
  "mov.f32 y0, %10;\n\t@!p0 sub.sat.f32 y0, 0f3F800000, y0;\n\t"                                       \
  "mov.f32 y1, %11;\n\t@!p1 sub.sat.f32 y1, 0f3F800000, y1;\n\t"                                       \
  "mov.b64 yy, {y0, y1};\n\tadd.rn.f32x2 yy, yy, %12;\n\tmov.b64 {y0, y1}, yy;\n\t"                     \
  "rcp.approx.ftz.f32 r0, y0;\n\trcp.approx.ftz.f32 r1, y1;\n\t"                                        \
  "and.b32 h0, r0, 0xffffe000;\n\tand.b32 h1, r1, 0xffffe000;\n\t"                                      \
  "mov.b64 rr, {r0, r1};\n\tmov.b64 hh, {h0, h1};\n\tsub.rn.f32x2 rr, rr, hh;\n\tmov.b64 {l0, l1}, rr;\n\t" \
  "cvt.rn.bf16x2.f32 c0, l0, h0;\n\tcvt.rn.bf16x2.f32 c1, l1, h1;\n\t"

// QSEP = false: the directly accumulated plane is the ones' (q == p): one predicate per entry.
template <bool QSEP>
__device__ __forceinline__ void h_pair(float t0, float t1, uint32_t bits, uint32_t qbits, uint32_t bit0, uint64_t eps2,
                                       float& x0, float& x1, uint32_t& hi0, uint32_t& hi1, uint32_t& c0, uint32_t& c1,
                                       uint32_t& qhi0, uint32_t& qhi1, uint32_t& qc0, uint32_t& qc1) {
  if constexpr (!QSEP) {
    asm("{\n\t.reg .pred p0, p1;\n\t.reg .b32 t, h0, h1, c0, c1;\n\t.reg .f32 y0, y1, r0, r1, l0, l1;\n\t.reg .b64 yy, rr, hh;\n\t"
        "and.b32 t, %13, %15;\n\tsetp.ne.b32 p0, t, 0;\n\t"
        "and.b32 t, %13, %16;\n\tsetp.ne.b32 p1, t, 0;\n\t" NBMF_H_PAIR_BODY
        "mov.b32 %2, h0;\n\tmov.b32 %3, h1;\n\tmov.b32 %4, c0;\n\tmov.b32 %5, c1;\n\t"
        "selp.b32 %6, h0, 0, p0;\n\tselp.b32 %7, h1, 0, p1;\n\tselp.b32 %8, c0, 0, p0;\n\tselp.b32 %9, c1, 0, p1;\n\t"
        "mov.f32 %0, y0;\n\tmov.f32 %1, y1;\n\t}\n"
        : "=f"(x0), "=f"(x1), "=r"(hi0), "=r"(hi1), "=r"(c0), "=r"(c1), "=r"(qhi0), "=r"(qhi1), "=r"(qc0), "=r"(qc1)
        : "f"(t0), "f"(t1), "l"(eps2), "r"(bits), "r"(qbits), "r"(bit0), "r"(bit0 << 1));
  } else {
    asm("{\n\t.reg .pred p0, p1, q0, q1;\n\t.reg .b32 t, h0, h1, c0, c1;\n\t.reg .f32 y0, y1, r0, r1, l0, l1;\n\t.reg .b64 yy, rr, hh;\n\t"
        "and.b32 t, %13, %15;\n\tsetp.ne.b32 p0, t, 0;\n\t"
        "and.b32 t, %13, %16;\n\tsetp.ne.b32 p1, t, 0;\n\t" NBMF_H_PAIR_BODY
        "and.b32 t, %14, %15;\n\tsetp.ne.b32 q0, t, 0;\n\t"
        "and.b32 t, %14, %16;\n\tsetp.ne.b32 q1, t, 0;\n\t"
        "mov.b32 %2, h0;\n\tmov.b32 %3, h1;\n\tmov.b32 %4, c0;\n\tmov.b32 %5, c1;\n\t"
        "selp.b32 %6, h0, 0, q0;\n\tselp.b32 %7, h1, 0, q1;\n\tselp.b32 %8, c0, 0, q0;\n\tselp.b32 %9, c1, 0, q1;\n\t"
        "mov.f32 %0, y0;\n\tmov.f32 %1, y1;\n\t}\n"
        : "=f"(x0), "=f"(x1), "=r"(hi0), "=r"(hi1), "=r"(c0), "=r"(c1), "=r"(qhi0), "=r"(qhi1), "=r"(qc0), "=r"(qc1)
        : "f"(t0), "f"(t1), "l"(eps2), "r"(bits), "r"(qbits), "r"(bit0), "r"(bit0 << 1));
  }
}
//   H pass, strict mask semantics (unobserved entries contribute nothing: _solver.py with the README / paper mask): the
//   unmasked plane is masked by the observation bit and x is replaced by 1 (log 1 = 0) where the entry is unobserved.
__device__ __forceinline__ void h_pair_strict(float t0, float t1, uint32_t bits, uint32_t qbits, uint32_t obits, uint32_t bit0,
                                              uint64_t eps2, float& x0, float& x1, uint32_t& hi0, uint32_t& hi1, uint32_t& c0,
                                              uint32_t& c1, uint32_t& qhi0, uint32_t& qhi1, uint32_t& qc0, uint32_t& qc1) {
  asm("{\n\t.reg .pred p0, p1, q0, q1;\n\t.reg .b32 t, h0, h1, c0, c1;\n\t.reg .f32 y0, y1, r0, r1, l0, l1;\n\t.reg .b64 yy, rr, hh;\n\t"
      "and.b32 t, %13, %15;\n\tsetp.ne.b32 p0, t, 0;\n\t"
      "and.b32 t, %13, %16;\n\tsetp.ne.b32 p1, t, 0;\n\t" NBMF_H_PAIR_BODY
      "and.b32 t, %14, %15;\n\tsetp.ne.b32 q0, t, 0;\n\t"
      "and.b32 t, %14, %16;\n\tsetp.ne.b32 q1, t, 0;\n\t"
      "selp.b32 %6, h0, 0, q0;\n\tselp.b32 %7, h1, 0, q1;\n\tselp.b32 %8, c0, 0, q0;\n\tselp.b32 %9, c1, 0, q1;\n\t"
      "and.b32 t, %17, %15;\n\tsetp.ne.b32 q0, t, 0;\n\t"
      "and.b32 t, %17, %16;\n\tsetp.ne.b32 q1, t, 0;\n\t"
      "selp.b32 %2, h0, 0, q0;\n\tselp.b32 %3, h1, 0, q1;\n\tselp.b32 %4, c0, 0, q0;\n\tselp.b32 %5, c1, 0, q1;\n\t"
      "selp.f32 %0, y0, 0f3F800000, q0;\n\tselp.f32 %1, y1, 0f3F800000, q1;\n\t}\n"
      : "=f"(x0), "=f"(x1), "=r"(hi0), "=r"(hi1), "=r"(c0), "=r"(c1), "=r"(qhi0), "=r"(qhi1), "=r"(qc0), "=r"(qc1)
      : "f"(t0), "f"(t1), "l"(eps2), "r"(bits), "r"(qbits), "r"(bit0), "r"(bit0 << 1), "r"(obits));
}
//   Loss-only H pass (the objective of the final factors, score / evaluate): only x is needed.
__device__ __forceinline__ float h_x(float theta, uint32_t bits, uint32_t bit, float eps) {
  float x;
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t.reg .f32 y;\n\t"
      "and.b32 t, %2, %3;\n\tsetp.ne.b32 p, t, 0;\n\t"
      "mov.f32 y, %1;\n\t@!p sub.sat.f32 y, 0f3F800000, y;\n\t"
      "add.f32 %0, y, %4;\n\t}\n"
      : "=f"(x) : "f"(theta), "r"(bits), "r"(bit), "f"(eps));
  return x;
}
__device__ __forceinline__ float h_x_strict(float theta, uint32_t bits, uint32_t obits, uint32_t bit, float eps) {
  float x;
  asm("{\n\t.reg .pred p, o;\n\t.reg .b32 t;\n\t.reg .f32 y;\n\t"
      "and.b32 t, %2, %4;\n\tsetp.ne.b32 p, t, 0;\n\t"
      "and.b32 t, %3, %4;\n\tsetp.ne.b32 o, t, 0;\n\t"
      "mov.f32 y, %1;\n\t@!p sub.sat.f32 y, 0f3F800000, y;\n\t"
      "add.f32 y, y, %5;\n\t"
      "selp.f32 %0, y, 0f3F800000, o;\n\t}\n"
      : "=f"(x) : "f"(theta), "r"(bits), "r"(obits), "r"(bit), "f"(eps));
  return x;
}
//   W pass, two entries: signed ratio s = 1/(theta + eps) on ones, -1/((1 - theta) + eps) on observed zeros, 0 on
//   unobserved entries; q accumulates the zeros' 1/x (= -s) with one predicated subtract; (hi, c) split as above.
__device__ __forceinline__ void w_pair(float t0, float t1, uint32_t pbits, uint32_t obits, uint32_t bit0, uint64_t eps2,
                                       float eps, float& q, uint32_t& hi0, uint32_t& hi1, uint32_t& c0, uint32_t& c1) {
  asm("{\n\t.reg .pred p0, p1, o0, o1;\n\t.reg .b32 t, h0, h1;\n\t.reg .f32 y0, y1, s0, s1, l0, l1;\n\t.reg .b64 yy, ss, hh;\n\t"
      "and.b32 t, %7, %9;\n\tsetp.ne.b32 p0, t, 0;\n\t"
      "and.b32 t, %7, %10;\n\tsetp.ne.b32 p1, t, 0;\n\t"
      "and.b32 t, %8, %9;\n\tsetp.ne.b32 o0, t, 0;\n\t"
      "and.b32 t, %8, %10;\n\tsetp.ne.b32 o1, t, 0;\n\t"
      "mov.b64 yy, {%5, %6};\n\tadd.rn.f32x2 yy, yy, %11;\n\tmov.b64 {y0, y1}, yy;\n\t"
      "@!p0 add.f32 y0, %5, 0fBF800000;\n\t@!p0 sub.f32 y0, y0, %12;\n\t"
      "@!p1 add.f32 y1, %6, 0fBF800000;\n\t@!p1 sub.f32 y1, y1, %12;\n\t"
      "rcp.approx.ftz.f32 y0, y0;\n\trcp.approx.ftz.f32 y1, y1;\n\t"
      "selp.f32 s0, y0, 0f00000000, o0;\n\tselp.f32 s1, y1, 0f00000000, o1;\n\t"
      "@!p0 sub.f32 %0, %0, s0;\n\t@!p1 sub.f32 %0, %0, s1;\n\t"
      "and.b32 h0, s0, 0xffffe000;\n\tand.b32 h1, s1, 0xffffe000;\n\t"
      "mov.b64 ss, {s0, s1};\n\tmov.b64 hh, {h0, h1};\n\tsub.rn.f32x2 ss, ss, hh;\n\tmov.b64 {l0, l1}, ss;\n\t"
      "cvt.rn.bf16x2.f32 %3, l0, h0;\n\tcvt.rn.bf16x2.f32 %4, l1, h1;\n\t"
      "mov.b32 %1, h0;\n\tmov.b32 %2, h1;\n\t}\n"
      : "+f"(q), "=r"(hi0), "=r"(hi1), "=r"(c0), "=r"(c1)
      : "f"(t0), "f"(t1), "r"(pbits), "r"(obits), "r"(bit0), "r"(bit0 << 1), "l"(eps2), "f"(eps));
}
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float prod4(float x0, float x1, float x2, float x3) {   // one packed multiply + one scalar
  uint64_t c;                                  // (x0, x1) and (x2, x3) are the register pairs the entry arithmetic left
  float lo, hi;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(pack2(x0, x1)), "l"(pack2(x2, x3)));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(c));
  return lo * hi;
}

// =====================================================================================
// H pass (TMEM lane = column j).  CTA = 128 columns j, streams 32-row blocks of W, half-block = 16 rows.
//   MMA1(hb): Theta^T[128 j x 16 i] = Ht[128 x KB k] . W[16 i x KB k]^T
//   SIMT:     bit i of the column-tiled plane Pc; r = 1/x; planes R = r and Rq = [q] r (hi, c each); fused NLL: one
//             MUFU.LG2 per product of four x
//   MMA2(hb): Q^T[128 j x KB k] += Rq[128 x 16 i] . W^T[KB k x 16 i]^T,  S^T += R . W^T
// q is the ones' bit, or -- per COLUMN, when the column's density of ones exceeds its mean H, i.e. when the ones' sum C
// is expected to dominate the zeros' sum D -- the zeros' bit, so that the smaller of C, D is accumulated directly and
// only the larger one comes from S - Q (forming the small one by subtraction loses log2(large / small) bits).
// TMEM: A hi 0..KT-1, bf16 [hi|lo] KT..2KT-1 | Theta[p] 16 each | R[p] 64 each (per 8 rows: Rq_hi R_hi Rq_c R_c)
//       | {Q, S}[set] NACC each
// =====================================================================================
template <int KB>
struct HTc {
  using C = TcCfg<KB>;
  static constexpr int STAGES = KB == 64 ? 4 : 8;
  static constexpr int LOOK = KB == 64 ? 3 : 2;                   // stages are produced this many own blocks ahead
  static constexpr int REG = C::SLABS * 4096;                     // one operand region of a stage (32 rows x KT / KT rows x 32)
  static constexpr int STAGE_BYTES = 4 * REG;
  static constexpr int NFT = KB == 64 ? 256 : 512;                // threads that own fp32 accumulators
  static constexpr int NV = KB == 64 ? 64 : KB;                   // accumulators per such thread
  static constexpr int OFF_ACC = STAGES * STAGE_BYTES;
  static constexpr int SMEM = OFF_ACC + NV * NFT * 4 + 1024;
  static constexpr int T_THETA = 2 * C::KT, T_R = T_THETA + 16 * C::P, T_ACC = T_R + 64 * C::P;
  static_assert(T_ACC + C::NSETS * 2 * C::NACC <= 512, "TMEM");
};

template <int KB, bool STRICT, bool CD>      // CD = false: loss-only pass (no ratio planes, no MMA2, no C / D output)
__global__ void __launch_bounds__(TcCfg<KB>::THREADS, 1) h_pass_tc_kernel(const HTcArgs a) {
  using namespace tc;
  using C = TcCfg<KB>;
  using L = HTc<KB>;
  constexpr int P = C::P, NSETS = C::NSETS, KT = C::KT, NACC = C::NACC, STAGES = L::STAGES;
  if (*NBMF_TSH(a.done)) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sAcc = reinterpret_cast<float*>(smem + L::OFF_ACC);        // [NV][NFT]
  __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES];
  __shared__ uint64_t bar_a, bar_theta[P], bar_tfree[P], bar_s[P], bar_rfree[P], bar_cd[NSETS], bar_flushed[NSETS];
  __shared__ uint32_t tmem_base_s;
  __shared__ double red_scratch[C::WARPS];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // provably warp-uniform
  const int64_t jb = (int64_t)blockIdx.x * 128;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * a.rows_per_split;
  const int64_t r1 = min(a.m, r0 + a.rows_per_split);
  const int nb = r1 > r0 ? (int)((r1 - r0 + 31) / 32) : 0;           // 32-row blocks; half-blocks 0 .. 2 nb - 1
  constexpr int FLUSHERS = KB == 64 ? 8 : 8;                         // warps that read one accumulator set

  if (warp == C::MMA1_WARP) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_a, C::SIMT_WARPS);
    for (int p = 0; p < P; ++p) {
      mbar_init(&bar_theta[p], 1); mbar_init(&bar_tfree[p], 4);
      mbar_init(&bar_s[p], 4); mbar_init(&bar_rfree[p], 1);
    }
    for (int s = 0; s < NSETS; ++s) { mbar_init(&bar_cd[s], 1); mbar_init(&bar_flushed[s], FLUSHERS); }
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tA = tb, tTheta = tb + L::T_THETA, tR = tb + L::T_R, tAcc = tb + L::T_ACC;
  constexpr uint32_t id1 = idesc_tf32(128, 16), id1b = idesc_bf16(128, 16);
  constexpr uint32_t id2 = idesc_tf32(128, NACC), id2b = idesc_bf16(128, NACC);

  double ll_total = 0.0;
  if (warp >= C::MMA1_WARP && warp < C::MMA2_WARP) {
    // ------------------------------------------------------------- MMA1 issuer e: 32-row blocks b = e, e + NM1, .., and
    // producer of their shared-memory stages: one bulk copy per block, LOOK own blocks ahead.  The stage of block
    // b + NM1 LOOK was last read by MMA2(b + NM1 LOOK - STAGES), long completed when MMA1(b) is issued.
    constexpr int NM1 = C::NM1;
    const int e = warp - C::MMA1_WARP;
    const bool leader = elect_one();
    const float* src = NBMF_TSH(a.Wf) + (size_t)(r0 >> 5) * (L::STAGE_BYTES / 4);
    auto produce = [&](int bp) {
      if (bp >= nb) return;
      const int s = bp % STAGES;
      if (bp >= STAGES) mbar_wait(&bar_empty[s], ((bp / STAGES) - 1) & 1);
      if (leader) {
        mbar_expect_tx(&bar_full[s], L::STAGE_BYTES);
        bulk_g2s(smem + s * L::STAGE_BYTES, src + (size_t)bp * (L::STAGE_BYTES / 4), L::STAGE_BYTES, &bar_full[s]);
      }
      __syncwarp();
    };
#pragma unroll
    for (int i = 0; i < L::LOOK; ++i) produce(e + NM1 * i);
    mbar_wait(&bar_a, 0);
    fence_after_sync();
    for (int b = e; b < nb; b += NM1) {
      const int s = b % STAGES;
      mbar_wait(&bar_full[s], (b / STAGES) & 1);
      const uint32_t st = smem_u32(smem + s * L::STAGE_BYTES);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int hb = 2 * b + h, p = hb % P, u = hb / P;
        if (u >= 1) mbar_wait(&bar_tfree[p], (u - 1) & 1);               // Theta(hb - P) sits in the SIMT registers
        fence_after_sync();
        if (leader) {
          const uint32_t tT = tTheta + 16 * p;
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks)                             // hi . hi, tf32
            mma_ts(tT, tA + 8 * ks, desc_kmajor_sw128(st + (ks >> 2) * 4096 + h * 2048) + 2 * (ks & 3), id1, ks > 0);
#pragma unroll
          for (int i = 0; i < C::KS; ++i) {                              // hi.lo + lo.hi, bf16: A [hi | lo], B [lo | hi]
            // KB = 16 lives in the K <= 32 formats: hi[0..15] is bf16 K step 0, lo[0..15] K step 2
            const int ks = KB == 16 ? 2 * i : i;
            mma_ts_bf16(tT, tA + KT + 8 * ks, desc_kmajor_sw128(st + L::REG + (ks >> 2) * 4096 + h * 2048) + 2 * (ks & 3), id1b, 1);
          }
          commit(&bar_theta[p]);
          if (!CD && h == 1) commit(&bar_empty[s]);                      // loss-only pass: MMA1 is the last reader
        }
        __syncwarp();
      }
      produce(b + NM1 * L::LOOK);
    }
  } else if (warp >= C::MMA2_WARP) {
    // ------------------------------------------------------------- MMA2 issuer of accumulator set `set`: its blocks
    // b = set, set + NSETS, .., both halves, in order (the accumulation order is fixed: deterministic results)
    const int set = warp - C::MMA2_WARP;
    const bool leader = elect_one();
    const uint32_t tQ = tAcc + 2 * NACC * set, tS = tQ + NACC;
    if (CD) {
      for (int b = set; b < nb; b += NSETS) {
        const int cb = b / NSETS;                                        // index among the set's own blocks
        const int s = b % STAGES;
        const uint32_t st = smem_u32(smem + s * L::STAGE_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int hb = 2 * b + h, p = hb % P, u = hb / P;
          const bool chain_start = (cb % kFlush) == 0 && h == 0;
          mbar_wait(&bar_s[p], u & 1);
          if (chain_start && cb > 0) mbar_wait(&bar_flushed[set], ((cb / kFlush) - 1) & 1);   // previous chain was read
          fence_after_sync();
          if (leader) {
            const uint64_t dTh = desc_kmajor_sw128(st + 2 * L::REG), dTc = desc_kmajor_sw128(st + 3 * L::REG);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int ks = 2 * h + kk;
              const uint32_t acc = (kk > 0 || !chain_start) ? 1u : 0u;
              const uint32_t tr = tR + 64 * p + 32 * kk;                 // Rq_hi +0, R_hi +8, Rq_c +16, R_c +24
              mma_ts(tQ, tr, dTh + 2 * ks, id2, acc);                    // hi . hi, tf32
              mma_ts(tS, tr + 8, dTh + 2 * ks, id2, acc);
              mma_ts_bf16(tQ, tr + 16, dTc + 2 * ks, id2b, 1);           // hi.lo + lo.hi, bf16, K = 16 = 8 rows x (hi, lo)
              mma_ts_bf16(tS, tr + 24, dTc + 2 * ks, id2b, 1);
            }
            commit(&bar_rfree[p]);
            if (h == 1) {
              commit(&bar_empty[s]);
              if (b + NSETS >= nb || ((cb + 1) % kFlush) == 0) commit(&bar_cd[set]);   // a chain ends here
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------- SIMT warps: pipeline p, lane quarter q; lane = column j,
    // 16 rows i per half-block
    const int q = warp & 3, p = warp >> 2;
    const int set = NSETS == 1 ? 0 : (p >> 1);
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int tl = q * 32 + lane;                                      // TMEM lane == column inside the CTA tile
    const int64_t col = jb + tl;
    const float eps = a.eps;
    const uint64_t eps2 = pack2(eps, eps);
    // resident A operand: this thread's column of H; the pipelines of a lane quarter share the K chunks of 8:
    // tf32 hi plane (KT columns) and the bf16 correction plane [hi (k = 0..KT-1) | lo], two elements per column
    float hsum = 0.f;
    const float* __restrict__ Hb = NBMF_TSH(a.H);
#pragma unroll
    for (int c = p; c < KT / 8; c += P) {
      uint32_t hi[8], bh[4], bl[4];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float x0 = Hb[(size_t)(8 * c + e) * a.ldh + col], x1 = Hb[(size_t)(8 * c + e + 1) * a.ldh + col];
        const float h0 = tf32_trunc(x0), h1 = tf32_trunc(x1);
        hi[e] = __float_as_uint(h0);
        hi[e + 1] = __float_as_uint(h1);
        bh[e >> 1] = bf16_bits(h0) | (bf16_bits(h1) << 16);
        bl[e >> 1] = bf16_bits(x0 - h0) | (bf16_bits(x1 - h1) << 16);
      }
      tmem_st8(tA + lane_off + 8 * c, hi);
      tmem_st4(tA + KT + lane_off + 4 * c, bh);
      tmem_st4(tA + KT + KT / 2 + lane_off + 4 * c, bl);
    }
    wait_st();
    fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_a);
    // which plane this COLUMN accumulates directly (see the header of the kernel): every warp that owns the column
    // decides the same way from the same inputs
    uint32_t fmask = 0u;
    if (CD && a.colcnt != nullptr && col < a.n) {
      for (int k = 0; k < a.k; ++k) hsum += Hb[(size_t)k * a.ldh + col];
      const float hbar = hsum / (float)a.k, d = (float)NBMF_TSH(a.colcnt)[col] / (float)a.m;
      if (d > hbar) fmask = 0xffffffffu;
    }
    const bool any_flip = __any_sync(0xffffffffu, fmask != 0u);
    // fp32 sums of this thread's accumulator slice live in shared memory: they are touched once per chain, registers
    // are what the hot loop is short of.  K <= 32: k = (NACC/2) h' .. of the set's Q and S, h' = p & 1; K <= 64:
    // pipeline 0 owns Q, pipeline 1 owns S, pipeline 2 nothing.
    const bool flusher = CD && (KB == 64 ? p < 2 : true);
    float* __restrict__ myacc = sAcc + (KB == 64 ? (p & 1) * 128 + tl : tid);
    if (flusher) {
#pragma unroll
      for (int e = 0; e < L::NV; ++e) myacc[e * L::NFT] = 0.f;
    }
    int flushed = 0;                                                   // chains of this set already flushed
    auto flush = [&]() {                                               // TMEM chain -> fp32 accumulators
      mbar_wait(&bar_cd[set], flushed & 1);
      fence_after_sync();
      if constexpr (KB == 64) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint32_t c[32];
          tmem_ld32(tAcc + NACC * (p & 1) + lane_off + 32 * part, c);
          wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) myacc[(32 * part + e) * L::NFT] += __uint_as_float(c[e]);
        }
      } else if constexpr (KB == 32) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {                         // Q, then S
          uint32_t c[16];
          tmem_ld16(tAcc + 2 * NACC * set + NACC * part + lane_off + 16 * (p & 1), c);
          wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) myacc[(16 * part + e) * L::NFT] += __uint_as_float(c[e]);
        }
      } else {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint32_t c[8];
          tmem_ld8(tAcc + 2 * NACC * set + NACC * part + lane_off + 8 * (p & 1), c);
          wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) myacc[(8 * part + e) * L::NFT] += __uint_as_float(c[e]);
        }
      }
      ++flushed;
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_flushed[set]);                   // the next chain may overwrite the accumulators
    };
    const uint32_t* __restrict__ pc = NBMF_TSH(a.Pc) + ((size_t)blockIdx.x * a.nrb + (size_t)(r0 >> 5)) * 128 + tl;
    const uint32_t* __restrict__ mc = nullptr;
    if constexpr (STRICT) mc = NBMF_TSH(a.Mc) + ((size_t)blockIdx.x * a.nrb + (size_t)(r0 >> 5)) * 128 + tl;
    const int nhb = 2 * nb;
    uint32_t word = p < nhb ? pc[(size_t)(p >> 1) * 128] : 0u, mword = 0u;
    if constexpr (STRICT) mword = p < nhb ? mc[(size_t)(p >> 1) * 128] : 0u;
    float ll = 0.f, ll_sum = 0.f, ll_c = 0.f;                          // fp64 is slow here: compensated fp32 sum
    // A probe of an mbarrier costs ~200 clk even when its phase completed long ago, so the Theta wait of the next
    // half-block is probed early and only checked where it is needed: the latency hides behind the arithmetic.
    bool ok_theta = false;
    int u = 0;
    for (int hb = p; hb < nhb; hb += P, ++u) {
      const int b = hb >> 1, h = hb & 1;
      const uint32_t bits = word >> (16 * h), obits = mword >> (16 * h);
      if (hb + P < nhb) {                                              // prefetch the next half-block's bits
        word = pc[(size_t)((hb + P) >> 1) * 128];
        if constexpr (STRICT) mword = mc[(size_t)((hb + P) >> 1) * 128];
      }
      uint32_t qbits = bits;
      if constexpr (STRICT) { if (fmask) qbits = obits & ~bits; } else { qbits = bits ^ fmask; }
      if (!ok_theta) mbar_wait(&bar_theta[p], u & 1);
      fence_after_sync();
      uint32_t v[16];
      {
        uint32_t v0[8], v1[8];
        tmem_ld8(tTheta + 16 * p + lane_off, v0);
        tmem_ld8(tTheta + 16 * p + lane_off + 8, v1);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[e] = v0[e]; v[8 + e] = v1[e]; }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tfree[p]);                       // Theta[p] may be overwritten by MMA1(hb + P)
      float llb = 0.f;
      auto arith = [&](auto qsep_tag) {                                // the 16 entries of this thread: ratios -> TMEM
        constexpr bool QSEP = decltype(qsep_tag)::value;
#pragma unroll
        for (int g = 0; g < 2; ++g) {                                  // 8 rows = one K step of MMA2
          uint32_t out[32];                                            // Rq_hi | R_hi | Rq_c | R_c, 8 columns each
          float xa = 1.f, xb = 1.f;
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float x0, x1;
            const float t0 = __uint_as_float(v[8 * g + e]), t1 = __uint_as_float(v[8 * g + e + 1]);
            const uint32_t bit0 = 1u << (8 * g + e);
            if constexpr (!CD) {
              x0 = STRICT ? h_x_strict(t0, bits, obits, bit0, eps) : h_x(t0, bits, bit0, eps);
              x1 = STRICT ? h_x_strict(t1, bits, obits, bit0 << 1, eps) : h_x(t1, bits, bit0 << 1, eps);
            } else if constexpr (STRICT) {
              h_pair_strict(t0, t1, bits, qbits, obits, bit0, eps2, x0, x1, out[8 + e], out[9 + e], out[24 + e], out[25 + e],
                            out[e], out[e + 1], out[16 + e], out[17 + e]);
            } else {
              h_pair<QSEP>(t0, t1, bits, qbits, bit0, eps2, x0, x1, out[8 + e], out[9 + e], out[24 + e], out[25 + e],
                           out[e], out[e + 1], out[16 + e], out[17 + e]);
            }
            // one log per product of four x: eps >= 1e-9 keeps four factors far above the underflow threshold
            if ((e & 2) == 0) { xa = x0; xb = x1; } else llb += logu_(prod4(xa, xb, x0, x1));
          }
          if (CD) {
            if (g == 0 && u >= 1) {                                    // MMA2(hb - P) must be done reading R[p]
              mbar_wait(&bar_rfree[p], (u - 1) & 1);
              fence_after_sync();
            }
            tmem_st32(tR + 64 * p + lane_off + 32 * g, out);
          }
        }
      };
      if (CD && !STRICT && any_flip) arith(std::true_type{}); else arith(std::false_type{});
      ok_theta = hb + P < nhb && mbar_try(smem_u32(&bar_theta[p]), (u + 1) & 1);
      if (CD) {
        if (flusher && ((b / NSETS) / kFlush) > flushed) flush();      // previous chain: its MMAs ended a while ago
        wait_st();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s[p]);
      ll += llb;
      if ((u & 15) == 15) {                                            // Kahan step every 16 own half-blocks
        const float y = ll - ll_c, t = ll_sum + y;
        ll_c = (t - ll_sum) - y;
        ll_sum = t;
        ll = 0.f;
      }
    }
    ll_total = ((double)ll_sum - (double)ll_c) + (double)ll;
    if (col >= a.n) ll_total = 0.0;
    if (CD) {
      const int nblk_set = nb > set ? (nb - set + NSETS - 1) / NSETS : 0;
      const int chains = (nblk_set + kFlush - 1) / kFlush;
      while (flusher && flushed < chains) flush();                     // the set's last chain(s)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * C::SIMT_WARPS) : "memory");   // the SIMT warps only
      float* __restrict__ base = NBMF_TSH(a.CD) + (size_t)(split * 2) * KT * a.ldh;     // C rows 0..KT-1 then D rows 0..KT-1
      if constexpr (KB == 64) {
        if (p == 0) {                                                  // Q from this thread, S from pipeline 1's (+128)
#pragma unroll 8
          for (int e = 0; e < 64; ++e) {
            const float qv = myacc[e * L::NFT], sv = myacc[e * L::NFT + 128];
            const float other = sv - qv;
            base[(size_t)e * a.ldh + col] = fmask ? other : qv;
            base[(size_t)(KT + e) * a.ldh + col] = fmask ? qv : other;
          }
        }
      } else {
        if (set == 0) {                                                // set 0 + set 1 (thread tid + 256), fixed order
          constexpr int HALF = NACC / 2;
          const int hp = p & 1;
#pragma unroll
          for (int e = 0; e < HALF; ++e) {                             // column j of row k: coalesced across the warp
            const float qv = myacc[e * L::NFT] + myacc[e * L::NFT + 256];
            const float sv = myacc[(HALF + e) * L::NFT] + myacc[(HALF + e) * L::NFT + 256];
            const float other = sv - qv;
            base[(size_t)(HALF * hp + e) * a.ldh + col] = fmask ? other : qv;
            base[(size_t)(KT + HALF * hp + e) * a.ldh + col] = fmask ? qv : other;
          }
          if constexpr (KB == 16) {                                    // rows 16..31 of the K <= 32 layout: padding
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              base[(size_t)(16 + 8 * hp + e) * a.ldh + col] = 0.f;
              base[(size_t)(KT + 16 + 8 * hp + e) * a.ldh + col] = 0.f;
            }
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == C::MMA1_WARP) tmem_dealloc(tb, 512);
  const double tot = block_sum<C::THREADS>(ll_total, red_scratch);
  if (tid == 0) NBMF_TSH(a.LL)[(size_t)split * gridDim.x + blockIdx.x] = tot * log_unit<float>();
}

// =====================================================================================
// W pass (TMEM lane = row i).  CTA = 128 rows i, streams 64-column blocks of H, half-block = 32 columns.
//   MMA1(hb): Theta'[128 i x 32 j] = W[128 x KB k] . Ht[32 j x KB k]^T
//   SIMT:     signed ratio s = 1/(+-x) on observed entries (= p - q), q-sum of the zeros' 1/x, s -> TMEM (hi, c)
//   MMA2(hb): G[128 i x KB k] += S[128 x 32 j] . H[KB k x 32 j]^T
// TMEM: A hi 0..KT-1, bf16 [hi|lo] KT..2KT-1 | Theta'[p] 32 each | S[p] 64 each (per 8 columns: S_hi S_c) | G[set] NACC each
// =====================================================================================
template <int KB>
struct WTc {
  using C = TcCfg<KB>;
  static constexpr int STAGES = KB == 64 ? 3 : 6;
  static constexpr int LOOK = KB == 64 ? 2 : 1;
  static constexpr int RW = C::SLABS * 8192;                      // Ht rows: 64 j x KT k, one 8 KB slab per 32 k
  static constexpr int RB = C::KT * 128;                          // one K-block of H: KT k rows x 32 j
  static constexpr int STAGE_BYTES = 4 * RW;                      // Ht hi | corr | H (2 K-blocks) hi | corr
  static constexpr int OFF_X = STAGES * STAGE_BYTES;              // set 1 -> set 0 accumulator exchange [NACC/2][256]
  static constexpr int OFF_Q = OFF_X + (KB == 64 ? 0 : (KB / 2) * 256 * 4);
  static constexpr int SMEM = OFF_Q + C::P * 128 * 4 + 1024;
  static constexpr int T_THETA = 2 * C::KT, T_S = T_THETA + 32 * C::P, T_G = T_S + 64 * C::P;
  static constexpr int NG = KB == 64 ? 32 : KB / 2;               // accumulator columns one SIMT thread owns
  static_assert(T_G + C::NSETS * C::NACC <= 512, "TMEM");
  static_assert(2 * RB == RW, "stage layout");
};

template <int KB>
__global__ void __launch_bounds__(TcCfg<KB>::THREADS, 1) w_pass_tc_kernel(const WTcArgs a) {
  using namespace tc;
  using C = TcCfg<KB>;
  using L = WTc<KB>;
  constexpr int P = C::P, NSETS = C::NSETS, KT = C::KT, NACC = C::NACC, STAGES = L::STAGES, NG = L::NG;
  if (*NBMF_TSH(a.done)) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sX = reinterpret_cast<float*>(smem + L::OFF_X);            // [NG][256 threads of set 1]
  float* sQ = reinterpret_cast<float*>(smem + L::OFF_Q);            // [P][128 lanes]
  __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES];
  __shared__ uint64_t bar_a, bar_theta[P], bar_tfree[P], bar_s[P], bar_sfree[P], bar_g[NSETS], bar_flushed[NSETS];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t ib = (int64_t)blockIdx.x * 128;
  const int64_t c0 = (int64_t)blockIdx.y * a.cols_per_split;
  const int64_t c1 = min(a.n, c0 + a.cols_per_split);
  const int nb = c1 > c0 ? (int)((c1 - c0 + 63) / 64) : 0;           // 64-column blocks; half-blocks 0 .. 2 nb - 1

  if (warp == C::MMA1_WARP) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_a, C::SIMT_WARPS);
    for (int p = 0; p < P; ++p) {
      mbar_init(&bar_theta[p], 1); mbar_init(&bar_tfree[p], 4);
      mbar_init(&bar_s[p], 4); mbar_init(&bar_sfree[p], 1);
    }
    for (int s = 0; s < NSETS; ++s) { mbar_init(&bar_g[s], 1); mbar_init(&bar_flushed[s], 8); }
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tA = tb, tTheta = tb + L::T_THETA, tS = tb + L::T_S, tG = tb + L::T_G;
  constexpr uint32_t id1 = idesc_tf32(128, 32), id1b = idesc_bf16(128, 32);
  constexpr uint32_t id2 = idesc_tf32(128, NACC), id2b = idesc_bf16(128, NACC);

  if (warp >= C::MMA1_WARP && warp < C::MMA2_WARP) {
    // ------------------------------------------------------------- MMA1 issuer e + producer of its stages (see the H pass)
    constexpr int NM1 = C::NM1;
    const int e = warp - C::MMA1_WARP;
    const bool leader = elect_one();
    const float* src = NBMF_TSH(a.Hf) + (size_t)(c0 >> 6) * (L::STAGE_BYTES / 4);
    auto produce = [&](int bp) {
      if (bp >= nb) return;
      const int s = bp % STAGES;
      if (bp >= STAGES) mbar_wait(&bar_empty[s], ((bp / STAGES) - 1) & 1);
      if (leader) {
        mbar_expect_tx(&bar_full[s], L::STAGE_BYTES);
        bulk_g2s(smem + s * L::STAGE_BYTES, src + (size_t)bp * (L::STAGE_BYTES / 4), L::STAGE_BYTES, &bar_full[s]);
      }
      __syncwarp();
    };
#pragma unroll
    for (int i = 0; i < L::LOOK; ++i) produce(e + NM1 * i);
    mbar_wait(&bar_a, 0);
    fence_after_sync();
    for (int b = e; b < nb; b += NM1) {
      const int s = b % STAGES;
      mbar_wait(&bar_full[s], (b / STAGES) & 1);
      const uint32_t st = smem_u32(smem + s * L::STAGE_BYTES);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int hb = 2 * b + h, p = hb % P, u = hb / P;
        if (u >= 1) mbar_wait(&bar_tfree[p], (u - 1) & 1);
        fence_after_sync();
        if (leader) {
          const uint32_t tT = tTheta + 32 * p;
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks)                             // hi . hi, tf32
            mma_ts(tT, tA + 8 * ks, desc_kmajor_sw128(st + (ks >> 2) * 8192 + h * 4096) + 2 * (ks & 3), id1, ks > 0);
#pragma unroll
          for (int i = 0; i < C::KS; ++i) {                              // hi.lo + lo.hi, bf16
            const int ks = KB == 16 ? 2 * i : i;
            mma_ts_bf16(tT, tA + KT + 8 * ks, desc_kmajor_sw128(st + L::RW + (ks >> 2) * 8192 + h * 4096) + 2 * (ks & 3), id1b, 1);
          }
          commit(&bar_theta[p]);
        }
        __syncwarp();
      }
      produce(b + NM1 * L::LOOK);
    }
  } else if (warp >= C::MMA2_WARP) {
    // ------------------------------------------------------------- MMA2 issuer of accumulator set `set`
    const int set = warp - C::MMA2_WARP;
    const bool leader = elect_one();
    const uint32_t tGa = tG + NACC * set;
    for (int b = set; b < nb; b += NSETS) {
      const int cb = b / NSETS;
      const int s = b % STAGES;
      const uint32_t st = smem_u32(smem + s * L::STAGE_BYTES);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int hb = 2 * b + h, p = hb % P, u = hb / P;
        const bool chain_start = (cb % kFlush) == 0 && h == 0;
        mbar_wait(&bar_s[p], u & 1);
        if (chain_start && cb > 0) mbar_wait(&bar_flushed[set], ((cb / kFlush) - 1) & 1);
        fence_after_sync();
        if (leader) {
          const uint64_t dBh = desc_kmajor_sw128(st + 2 * L::RW + h * L::RB), dBc = desc_kmajor_sw128(st + 3 * L::RW + h * L::RB);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {                               // 8 columns j per K step
            const uint32_t ts = tS + 64 * p + 16 * kk;                   // S_hi +0, S_c +8
            mma_ts(tGa, ts, dBh + 2 * kk, id2, (kk > 0 || !chain_start) ? 1u : 0u);   // hi . hi, tf32
            mma_ts_bf16(tGa, ts + 8, dBc + 2 * kk, id2b, 1);             // hi.lo + lo.hi, bf16
          }
          commit(&bar_sfree[p]);
          if (h == 1) {
            commit(&bar_empty[s]);
            if (b + NSETS >= nb || ((cb + 1) % kFlush) == 0) commit(&bar_g[set]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------- SIMT warps: pipeline p, lane quarter q; lane = row i,
    // 32 columns j per half-block
    const int q = warp & 3, p = warp >> 2;
    const int set = NSETS == 1 ? 0 : (p >> 1);
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int tl = q * 32 + lane;
    const int64_t row = ib + tl;
    const float eps = a.eps;
    const uint64_t eps2 = pack2(eps, eps);
    // resident A operand: this thread's row of W (zero beyond m), K chunks of 8 shared by the pipelines of the quarter
#pragma unroll
    for (int c = p; c < KT / 8; c += P) {
      float x[8];
      if (row < a.m) {
        const float* __restrict__ Wb = NBMF_TSH(a.W);
        const float4 x0 = *reinterpret_cast<const float4*>(Wb + (size_t)row * KT + 8 * c);
        const float4 x1 = *reinterpret_cast<const float4*>(Wb + (size_t)row * KT + 8 * c + 4);
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = 0.f;
      }
      uint32_t hi[8], bh[4], bl[4];                                   // see the H pass: tf32 hi plane + bf16 [hi | lo]
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float h0 = tf32_trunc(x[e]), h1 = tf32_trunc(x[e + 1]);
        hi[e] = __float_as_uint(h0);
        hi[e + 1] = __float_as_uint(h1);
        bh[e >> 1] = bf16_bits(h0) | (bf16_bits(h1) << 16);
        bl[e >> 1] = bf16_bits(x[e] - h0) | (bf16_bits(x[e + 1] - h1) << 16);
      }
      tmem_st8(tA + lane_off + 8 * c, hi);
      tmem_st4(tA + KT + lane_off + 4 * c, bh);
      tmem_st4(tA + KT + KT / 2 + lane_off + 4 * c, bl);
    }
    wait_st();
    fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_a);
    // fp32 register accumulators: K <= 32: k = NG h' .. of the set's G, h' = p & 1; K <= 64: pipelines 0, 1 own 32 columns each
    const bool flusher = KB == 64 ? p < 2 : true;
    const int gcol = NG * (p & 1);
    float accG[NG];
#pragma unroll
    for (int e = 0; e < NG; ++e) accG[e] = 0.f;
    int flushed = 0;
    auto flush = [&]() {                                               // TMEM chain -> fp32 register accumulators
      mbar_wait(&bar_g[set], flushed & 1);
      fence_after_sync();
      if constexpr (NG == 32) {
        uint32_t c[32];
        tmem_ld32(tG + NACC * set + lane_off + gcol, c);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) accG[e] += __uint_as_float(c[e]);
      } else if constexpr (NG == 16) {
        uint32_t c[16];
        tmem_ld16(tG + NACC * set + lane_off + gcol, c);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) accG[e] += __uint_as_float(c[e]);
      } else {
        uint32_t c[8];
        tmem_ld8(tG + NACC * set + lane_off + gcol, c);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) accG[e] += __uint_as_float(c[e]);
      }
      ++flushed;
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_flushed[set]);
    };
    const int nhb = 2 * nb;
    const uint2* __restrict__ pm = NBMF_TSH(a.PM) + ((size_t)blockIdx.x * a.wpr + (size_t)(c0 >> 5)) * 128 + tl;
    uint2 word = p < nhb ? pm[(size_t)p * 128] : make_uint2(0u, 0u);
    float qsum = 0.f;
    bool ok_theta = false;                                             // early barrier probes, see the H pass
    int u = 0;
    for (int hb = p; hb < nhb; hb += P, ++u) {
      const int b = hb >> 1;
      const uint2 bits = word;
      if (hb + P < nhb) word = pm[(size_t)(hb + P) * 128];
      if (!ok_theta) mbar_wait(&bar_theta[p], u & 1);
      fence_after_sync();
      float qb = 0.f;                                                  // this half-block's sum over the observed zeros of 1/x
      bool ok_sfree = true;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t v[16];
        tmem_ld16(tTheta + 32 * p + lane_off + 16 * g, v);
        if (g == 0 && u >= 1) ok_sfree = mbar_try(smem_u32(&bar_sfree[p]), (u - 1) & 1);
        wait_ld();
        if (g == 1) {                                                  // Theta'[p] may be overwritten by MMA1(hb + P)
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_tfree[p]);
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {                                  // 8 entries = one K step of MMA2: S_hi | S_c
          uint32_t out[16];
#pragma unroll
          for (int e = 0; e < 8; e += 2)
            w_pair(__uint_as_float(v[8 * w + e]), __uint_as_float(v[8 * w + e + 1]), bits.x, bits.y, 1u << (16 * g + 8 * w + e),
                   eps2, eps, qb, out[e], out[e + 1], out[8 + e], out[9 + e]);
          if (g == 0 && w == 0) {                                      // MMA2(hb - P) must be done reading S[p]
            if (!ok_sfree) mbar_wait(&bar_sfree[p], (u - 1) & 1);
            fence_after_sync();
          }
          tmem_st16(tS + 64 * p + lane_off + 16 * (2 * g + w), out);
        }
      }
      ok_theta = hb + P < nhb && mbar_try(smem_u32(&bar_theta[p]), (u + 1) & 1);
      if (flusher && ((b / NSETS) / kFlush) > flushed) flush();        // previous chain: its MMAs ended a while ago
      wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s[p]);
      qsum += qb;
    }
    {
      const int nblk_set = nb > set ? (nb - set + NSETS - 1) / NSETS : 0;
      const int chains = (nblk_set + kFlush - 1) / kFlush;
      while (flusher && flushed < chains) flush();
    }
    sQ[p * 128 + tl] = qsum;
    if constexpr (KB != 64) {
      const int t = (p & 1) * 128 + tl;                                // same (lane, k range) in both sets
      if (set == 1) {
#pragma unroll
        for (int e = 0; e < NG; ++e) sX[e * 256 + t] = accG[e];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * C::SIMT_WARPS) : "memory");   // the SIMT warps only
      if (set == 0 && row < a.m) {                                     // set 0 + set 1, fixed order
        float* __restrict__ Gg = NBMF_TSH(a.G) + ((size_t)blockIdx.y * a.m + row) * KT + gcol;
#pragma unroll
        for (int e = 0; e < NG; e += 4)
          *reinterpret_cast<float4*>(Gg + e) =
              make_float4(accG[e] + sX[e * 256 + t], accG[e + 1] + sX[(e + 1) * 256 + t],
                          accG[e + 2] + sX[(e + 2) * 256 + t], accG[e + 3] + sX[(e + 3) * 256 + t]);
        if constexpr (KB == 16) {                                      // columns 16..31 of the K <= 32 layout: padding
#pragma unroll
          for (int e = 0; e < 8; e += 4) *reinterpret_cast<float4*>(Gg + 16 - gcol + 8 * (p & 1) + e) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (p == 0)
          NBMF_TSH(a.Q)[(size_t)blockIdx.y * a.m + row] = ((sQ[tl] + sQ[128 + tl]) + sQ[256 + tl]) + sQ[384 + tl];
      }
    } else {
      asm volatile("bar.sync 1, %0;" ::"n"(32 * C::SIMT_WARPS) : "memory");
      if (p < 2 && row < a.m) {                                        // pipelines 0, 1 hold G columns 0..31, 32..63
        float* __restrict__ Gg = NBMF_TSH(a.G) + ((size_t)blockIdx.y * a.m + row) * KT + gcol;
#pragma unroll
        for (int e = 0; e < NG; e += 4)
          *reinterpret_cast<float4*>(Gg + e) = make_float4(accG[e], accG[e + 1], accG[e + 2], accG[e + 3]);
        if (p == 0) NBMF_TSH(a.Q)[(size_t)blockIdx.y * a.m + row] = (sQ[tl] + sQ[128 + tl]) + sQ[256 + tl];
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == C::MMA1_WARP) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------------------------ launchers
template <int KB>
inline void launch_w_pass_tc_kb(const WTcArgs& a, int nsplit, int batch_n, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set{0};
  ensure_dynamic_smem(w_pass_tc_kernel<KB>, WTc<KB>::SMEM, attr_set);
  dim3 grid((unsigned)((a.m + 127) / 128), (unsigned)nsplit, (unsigned)batch_n);
  w_pass_tc_kernel<KB><<<grid, TcCfg<KB>::THREADS, WTc<KB>::SMEM, st>>>(a);
}
template <int KB>
inline void launch_h_pass_tc_kb(const HTcArgs& a, int nsplit, int batch_n, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set[4];
  ensure_dynamic_smem(h_pass_tc_kernel<KB, false, true>, HTc<KB>::SMEM, attr_set[0]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, true, true>, HTc<KB>::SMEM, attr_set[1]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, false, false>, HTc<KB>::SMEM, attr_set[2]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, true, false>, HTc<KB>::SMEM, attr_set[3]);
  dim3 grid((unsigned)((a.n + 127) / 128), (unsigned)nsplit, (unsigned)batch_n);
  constexpr int T = TcCfg<KB>::THREADS, S = HTc<KB>::SMEM;
  if (a.compute_cd) {
    if (a.Mc) h_pass_tc_kernel<KB, true, true><<<grid, T, S, st>>>(a);
    else h_pass_tc_kernel<KB, false, true><<<grid, T, S, st>>>(a);
  } else {
    if (a.Mc) h_pass_tc_kernel<KB, true, false><<<grid, T, S, st>>>(a);
    else h_pass_tc_kernel<KB, false, false><<<grid, T, S, st>>>(a);
  }
}

}  // namespace k64
}  // namespace nbmf

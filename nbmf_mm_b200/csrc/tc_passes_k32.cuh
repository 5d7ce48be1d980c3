// Tensor-core (tcgen05 / TMEM) versions of the two pass kernels for K <= 32, fp32, bit-packed V (32 < K <= 64: tc_passes_k64.cuh).
// Template KB = 16 | 32: K extent the MMAs cover.  KB = 16 uses the same operand formats and skips the MMA K steps that
// are all padding and half of MMA2's N.
//
// Same math as passes.cuh, restructured like attention on Blackwell (S = QK^T -> P -> O = PV):
//   MMA1  Theta tile = resident factor tile [TMEM] . streamed factor block [smem]   (tcgen05.mma, TS form)
//   SIMT  tcgen05.ld Theta -> masked ratio in registers -> tcgen05.st back to TMEM
//   MMA2  accumulator (+)= ratio [TMEM] . streamed factor block [smem]               (tcgen05.mma, TS form)
// Both MMAs take their A operand from TMEM: an SS-form MMA with a 128-row A tile re-reads 4 KB of shared
// memory per K step and is bound by the 128 B/clk shared-memory port (40 clk per N=32 MMA, measured),
// the TS form runs at the tensor pipe's 16 clk (tools/tc_bench.cu).
//
// fp32 accuracy comes from a 3-term TF32 split: x = hi + lo with hi = x truncated to tf32 (what the
// tensor core reads), lo = x - hi, and a.b ~ hi.hi + hi.lo + lo.hi (~1e-6 relative, tools/tc_probe.cu).
// The streamed operand blocks are pre-split and pre-swizzled in global memory by format_factors.cu, so
// each pipeline stage is ONE 1-D bulk copy (cp.async.bulk + mbarrier complete_tx): no tensor maps.
//
// CTA = 20 warps (640 threads, 96 registers each), one CTA per SM (it owns all 512 TMEM columns).  The streamed
// blocks are dealt to two independent pipelines ("groups"): group g owns blocks b = g, g+2, .., its own Theta and
// ratio regions and its own accumulators in TMEM, eight SIMT warps and two issuing warps.
//   warps 0..15   SIMT.  Warp w works on TMEM lane quarter (w & 3) -- the hardware restricts a warp to
//                 lanes 32*(w % 4).. --, group g = (w >> 2) & 1 and half h = w >> 3 of the block's columns.
//   warps 16, 17  MMA1 issuer of group 0 / 1, which also produces the group's shared-memory stages (one bulk copy
//                 per block, a few blocks ahead);  warps 18, 19  MMA2 issuer of group 0 / 1.  The whole warp runs
//                 the loop (uniform control flow, operands in uniform registers); one elected lane executes
//                 the tcgen05.mma / commit / bulk-copy instructions.  One warp cannot issue everything: its barrier
//                 waits and commits are serial and cost ~200 clk each (measured with tools/tc_trace.cu).
// The tensor core's fp32 accumulate truncates, so an unbroken chain of 3e4 accumulations drifts by ~1e-4
// relative (measured); each group's TMEM accumulators are flushed into fp32 registers by the group's own
// SIMT warps every kFlush blocks, just before they release the first block of the next chain.
#pragma once
#include "args.h"
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_args.h"

#include <type_traits>

// pointer `ptr` of the kernel arguments `a` for fit blockIdx.z of a batch (workspaces batch_stride bytes apart)
#define NBMF_TSH(ptr) batch_shift(ptr, (size_t)blockIdx.z * (size_t)a.batch_stride)

namespace nbmf {
namespace k32 {

// Optional cycle trace of CTA (0,0) (tools/tc_trace.cu defines TC_TRACE): event e of warp slot w at block b.
#ifdef TC_TRACE
__device__ long long* g_tc_trace = nullptr;       // [4 warp slots][TC_TRACE_BLOCKS][8 events]
#define TC_TRACE_BLOCKS 256
#define TC_EV(slot, b, e)                                                                                     \
  do {                                                                                                        \
    if (g_tc_trace && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (b) < TC_TRACE_BLOCKS) \
      g_tc_trace[((slot) * TC_TRACE_BLOCKS + (b)) * 8 + (e)] = clock64();                                     \
  } while (0)
#else
#define TC_EV(slot, b, e) do { } while (0)
#endif

// Per-entry ratio arithmetic of the tensor kernels, written as PTX so that ONE predicate per entry (the data
// bit) drives both the choice of x and the masking of the outputs: ptxas otherwise rebuilds a 32-bit mask
// per entry with two shifts, and the ALU pipe (shifts, logic, selects) is the busiest SIMT pipe here.
//   H pass: x = (p ? theta : 1 - theta) + eps (1 - theta saturated to [0, 1]: a Theta that rounding pushed past 1
//   must not turn x negative; free), r = 1/x, hi = tf32(r), c = bf16x2(hi, r - hi) in one 32-bit
//   word (low half = hi: the K order of the bf16 correction MMA), and the copies of hi and c masked by p.
__device__ __forceinline__ void h_entry(float theta, uint32_t bits, uint32_t bit, float eps, float& x, uint32_t& hi,
                                        uint32_t& c, uint32_t& phi, uint32_t& pc) {
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t.reg .f32 y, r, h, l;\n\t"
      "and.b32 t, %6, %7;\n\tsetp.ne.b32 p, t, 0;\n\t"
      "mov.f32 y, %5;\n\t@!p sub.sat.f32 y, 0f3F800000, y;\n\t"
      "add.f32 y, y, %8;\n\t"
      "rcp.approx.ftz.f32 r, y;\n\t"
      "and.b32 h, r, 0xffffe000;\n\t"
      "sub.f32 l, r, h;\n\t"
      "cvt.rn.bf16x2.f32 %2, l, h;\n\t"
      "mov.b32 %1, h;\n\t"
      "selp.b32 %3, %1, 0, p;\n\t"
      "selp.b32 %4, %2, 0, p;\n\t"
      "mov.f32 %0, y;\n\t}\n"
      : "=f"(x), "=r"(hi), "=r"(c), "=r"(phi), "=r"(pc)
      : "f"(theta), "r"(bits), "r"(bit), "f"(eps));
}
//   The same with the directly accumulated plane selected by a second bit word (`qbits`): the zeros' plane of a column
//   whose ones dominate (see h_pass_tc_kernel).  Costs a second predicate per entry; only warps that hold such a column
//   take this path.
__device__ __forceinline__ void h_entry_q(float theta, uint32_t bits, uint32_t qbits, uint32_t bit, float eps, float& x,
                                          uint32_t& hi, uint32_t& c, uint32_t& phi, uint32_t& pc) {
  asm("{\n\t.reg .pred p, q;\n\t.reg .b32 t;\n\t.reg .f32 y, r, h, l;\n\t"
      "and.b32 t, %6, %8;\n\tsetp.ne.b32 p, t, 0;\n\t"
      "and.b32 t, %7, %8;\n\tsetp.ne.b32 q, t, 0;\n\t"
      "mov.f32 y, %5;\n\t@!p sub.sat.f32 y, 0f3F800000, y;\n\t"
      "add.f32 y, y, %9;\n\t"
      "rcp.approx.ftz.f32 r, y;\n\t"
      "and.b32 h, r, 0xffffe000;\n\t"
      "sub.f32 l, r, h;\n\t"
      "cvt.rn.bf16x2.f32 %2, l, h;\n\t"
      "mov.b32 %1, h;\n\t"
      "selp.b32 %3, %1, 0, q;\n\t"
      "selp.b32 %4, %2, 0, q;\n\t"
      "mov.f32 %0, y;\n\t}\n"
      : "=f"(x), "=r"(hi), "=r"(c), "=r"(phi), "=r"(pc)
      : "f"(theta), "r"(bits), "r"(qbits), "r"(bit), "f"(eps));
}
//   H pass, strict mask semantics (unobserved entries contribute nothing: _solver.py with the README/paper mask):
//   additionally the unmasked outputs are zeroed and x is replaced by 1 (log 1 = 0) where the entry is unobserved.
__device__ __forceinline__ void h_entry_strict(float theta, uint32_t bits, uint32_t obits, uint32_t qbits, uint32_t bit, float eps,
                                               float& x, uint32_t& hi, uint32_t& c, uint32_t& phi, uint32_t& pc) {
  asm("{\n\t.reg .pred p, o, q;\n\t.reg .b32 t, hh, cc;\n\t.reg .f32 y, r, h, l;\n\t"
      "and.b32 t, %6, %8;\n\tsetp.ne.b32 p, t, 0;\n\t"
      "and.b32 t, %7, %8;\n\tsetp.ne.b32 o, t, 0;\n\t"
      "and.b32 t, %10, %8;\n\tsetp.ne.b32 q, t, 0;\n\t"
      "mov.f32 y, %5;\n\t@!p sub.sat.f32 y, 0f3F800000, y;\n\t"
      "add.f32 y, y, %9;\n\t"
      "rcp.approx.ftz.f32 r, y;\n\t"
      "and.b32 h, r, 0xffffe000;\n\t"
      "sub.f32 l, r, h;\n\t"
      "cvt.rn.bf16x2.f32 cc, l, h;\n\t"
      "mov.b32 hh, h;\n\t"
      "selp.b32 %3, hh, 0, q;\n\t"
      "selp.b32 %4, cc, 0, q;\n\t"
      "selp.b32 %1, hh, 0, o;\n\t"
      "selp.b32 %2, cc, 0, o;\n\t"
      "selp.f32 %0, y, 0f3F800000, o;\n\t}\n"
      : "=f"(x), "=r"(hi), "=r"(c), "=r"(phi), "=r"(pc)
      : "f"(theta), "r"(bits), "r"(obits), "r"(bit), "f"(eps), "r"(qbits));
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
//   W pass: signed ratio s = 1/(theta + eps) on ones, -1/((1 - theta) + eps) on observed zeros, 0 on unobserved
//   entries; q accumulates the zeros' 1/x (= -s) with one predicated subtract.  `te` is Theta + eps as MMA1 delivers it
//   (its H operand is formatted from H + eps and W's rows sum to one: format_h_kernel), so the ones' x costs nothing and
//   the zeros' x = (te - 1) - 2 eps = -((1 - Theta) + eps) two predicated FADDs; eps2 = 2 eps.
__device__ __forceinline__ void w_entry(float te, uint32_t pbits, uint32_t obits, uint32_t bit, float eps2, float& s,
                                        float& q) {
  asm("{\n\t.reg .pred p, o;\n\t.reg .b32 t;\n\t.reg .f32 y;\n\t"
      "and.b32 t, %3, %5;\n\tsetp.ne.b32 p, t, 0;\n\t"
      "and.b32 t, %4, %5;\n\tsetp.ne.b32 o, t, 0;\n\t"
      "mov.f32 y, %2;\n\t"
      "@!p add.f32 y, %2, 0fBF800000;\n\t"
      "@!p sub.f32 y, y, %6;\n\t"
      "rcp.approx.ftz.f32 y, y;\n\t"
      "selp.f32 %0, y, 0f00000000, o;\n\t"
      "@!p sub.f32 %1, %1, %0;\n\t}\n"
      : "=f"(s), "+f"(q)
      : "f"(te), "r"(pbits), "r"(obits), "r"(bit), "f"(eps2));
}

constexpr int TC_SIMT_WARPS = 16;
constexpr int TC_MMA1_WARP = 16;                        // + group
constexpr int TC_MMA2_WARP = 18;                        // + group
constexpr int TC_THREADS = 20 * 32;                     // 640 threads: 96 registers each
constexpr int kFlush = 8;                                // own blocks per TMEM accumulation chain

// =====================================================================================
// H pass (TMEM lane = column j).  CTA = 128 columns j, streams 32-row blocks of W.
//   MMA1: Theta^T[128 j x 32 i] = Ht[128 x 32 k] . W[32 i x 32 k]^T
//   SIMT: bit i of the column-tiled plane Pc; r = 1/x; planes Rp = [p] r and R = r (hi, lo each);
//         fused NLL: one MUFU.LG2 per product of four x
//   MMA2: C^T[128 j x 32 k] += Rp[128 x 32 i] . W^T[32 k x 32 i]^T,  S^T += R . W^T;  D = S - C at the end
// TMEM: A hi 0..31, bf16 [hi|lo] 32..63 | Theta[g] 64..127 | R[g] 128..383 (per 8 rows: Rp_hi R_hi Rp_c R_c)
//       | {C, S}[g] 384..511
// =====================================================================================
constexpr int HTC_STAGES = 8;
constexpr int HTC_LOOK = 2;                              // stages are produced this many own blocks ahead
constexpr int HTC_STAGE_BYTES = 16384;
constexpr int HTC_OFF_ACC = HTC_STAGES * HTC_STAGE_BYTES;        // fp32 accumulators [32][512 SIMT threads]
constexpr int HTC_SMEM = HTC_OFF_ACC + 32 * 512 * 4 + 1024;

// Which plane a COLUMN accumulates directly: Rq is the ones' plane (C direct, D = S - C) unless the column's density of
// ones exceeds its mean H -- then the ones' sum C is expected to dominate the zeros' sum D (C / D ~ odds(density) /
// odds(mean Theta)), Rq becomes the zeros' plane and C = S - D: forming the SMALL one by subtraction would lose
// log2(large / small) bits.  Every warp that owns the column decides alike (same inputs); the choice is per TMEM lane.
// The decision is taken per H pass by flip_cols_kernel (below); FLIP = false is the instantiation for the common case
// that no column flips (one predicate per entry: the round-1 inner loop, untouched), FLIP = true the one with a second
// predicate per entry.  Both are launched, the one that does not match the device-side flag returns at once.
template <int KB, bool STRICT, bool CD, bool FLIP>   // CD = false: loss-only pass (no ratio planes, no MMA2, no C / D output)
__global__ void __launch_bounds__(TC_THREADS, 1) h_pass_tc_kernel(const HTcArgs a) {
  static_assert(KB == 16 || KB == 32, "KB");
  static_assert(CD || !FLIP, "the loss-only pass has no planes to choose");
  if (CD && a.flip_any != nullptr && (*NBMF_TSH(a.flip_any) != 0) != FLIP) return;
  using namespace tc;
  if (*NBMF_TSH(a.done)) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sAcc = reinterpret_cast<float*>(smem + HTC_OFF_ACC);      // thread t: 16 C then 16 S sums at [e * 512 + t]
  __shared__ uint64_t bar_full[HTC_STAGES], bar_empty[HTC_STAGES];
  __shared__ uint64_t bar_a, bar_theta[2], bar_tfree[2], bar_s[2], bar_rfree[2][2], bar_cd[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double red_scratch[TC_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // provably warp-uniform
  const int64_t jb = (int64_t)blockIdx.x * 128;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * a.rows_per_split;
  const int64_t r1 = min(a.m, r0 + a.rows_per_split);
  const int nb = r1 > r0 ? (int)((r1 - r0 + 31) / 32) : 0;
  constexpr bool cd = CD;

  if (warp == TC_MMA1_WARP) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < HTC_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_a, TC_SIMT_WARPS);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bar_theta[g], 1); mbar_init(&bar_tfree[g], TC_SIMT_WARPS / 2);
      mbar_init(&bar_s[g], TC_SIMT_WARPS / 2);
      mbar_init(&bar_rfree[g][0], 1); mbar_init(&bar_rfree[g][1], 1); mbar_init(&bar_cd[g], 1);
    }
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tA = tb, tTheta = tb + 64, tR = tb + 128, tAcc = tb + 384;
  constexpr uint32_t id = idesc_tf32(128, 32), idb = idesc_bf16(128, 32);       // MMA1: N = 32 rows of the block
  constexpr uint32_t id2 = idesc_tf32(128, KB), id2b = idesc_bf16(128, KB);     // MMA2: N = KB components
  constexpr int KS = KB / 8;                                                    // K steps of MMA1 per chain

  double ll_total = 0.0;
  if (warp == TC_MMA1_WARP || warp == TC_MMA1_WARP + 1) {
    // ------------------------------------------------------------- MMA1 issuer of group g, and producer of the
    // group's shared-memory stages: one 16 KB bulk copy per block, HTC_LOOK own blocks ahead.  The stage of block
    // b + 2 LOOK was last read by MMA2(b + 2 LOOK - STAGES) = MMA2(b - 4), which has completed by the time
    // MMA1(b) is issued (the SIMT warps are already working on block b - 2), so the wait does not stall the issue.
    const int g = warp - TC_MMA1_WARP;
    const bool leader = elect_one();
    const uint32_t tT = tTheta + 32 * g;
    const float* src = NBMF_TSH(a.Wf) + (size_t)(r0 >> 5) * 4096;
    auto produce = [&](int bp) {
      if (bp >= nb) return;
      const int s = bp % HTC_STAGES;
      if (bp >= HTC_STAGES) mbar_wait(&bar_empty[s], ((bp / HTC_STAGES) - 1) & 1);
      if (leader) {
        mbar_expect_tx(&bar_full[s], HTC_STAGE_BYTES);
        bulk_g2s(smem + s * HTC_STAGE_BYTES, src + (size_t)bp * 4096, HTC_STAGE_BYTES, &bar_full[s]);
      }
      __syncwarp();
    };
#pragma unroll
    for (int i = 0; i < HTC_LOOK; ++i) produce(g + 2 * i);
    mbar_wait(&bar_a, 0);
    fence_after_sync();
    for (int b = g; b < nb; b += 2) {
      const int s = b % HTC_STAGES;
      mbar_wait(&bar_full[s], (b / HTC_STAGES) & 1);
      if (b >= 2) mbar_wait(&bar_tfree[g], ((b >> 1) - 1) & 1);        // Theta(b-2) sits in the SIMT registers
      fence_after_sync();
      TC_EV(0, b, 0);
      if (leader) {
        const uint32_t st = smem_u32(smem + s * HTC_STAGE_BYTES);
        const uint64_t dWh = desc_kmajor_sw128(st), dWc = desc_kmajor_sw128(st + 4096);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma_ts(tT, tA + 8 * ks, dWh + 2 * ks, id, ks > 0);           // hi . hi, tf32
#pragma unroll
        for (int i = 0; i < KS; ++i) {                                 // hi.lo + lo.hi, bf16: A [hi | lo], B [lo | hi];
          const int ks = KB == 16 ? 2 * i : i;                         // KB = 16: hi[0..15] is K step 0, lo[0..15] K step 2
          mma_ts_bf16(tT, tA + 32 + 8 * ks, dWc + 2 * ks, idb, 1);
        }
        commit(&bar_theta[g]);
        if (!cd) commit(&bar_empty[s]);                                // loss-only pass: MMA1 is the last reader
      }
      __syncwarp();
      TC_EV(0, b, 1);
      produce(b + 2 * HTC_LOOK);
    }
  } else if (warp == TC_MMA2_WARP || warp == TC_MMA2_WARP + 1) {
    // ------------------------------------------------------------- MMA2 issuer of group g
    const int g = warp - TC_MMA2_WARP;
    const bool leader = elect_one();
    const uint32_t tRb = tR + 128 * g, tC = tAcc + 64 * g, tS = tC + 32;
    if (cd) {
      for (int b = g; b < nb; b += 2) {
        const int ob = b >> 1;                                         // index among the group's own blocks
        const bool chain_start = (ob % kFlush) == 0;
        mbar_wait(&bar_s[g], ob & 1);
        fence_after_sync();
        TC_EV(0, b, 2);
        if (leader) {
          const int s = b % HTC_STAGES;
          const uint32_t st = smem_u32(smem + s * HTC_STAGE_BYTES);
          const uint64_t dTh = desc_kmajor_sw128(st + 8192), dTc = desc_kmajor_sw128(st + 12288);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (ks > 0 || !chain_start) ? 1u : 0u;
            const uint32_t tr = tRb + 32 * ks;                         // Rp_hi +0, R_hi +8, Rp_c +16, R_c +24
            mma_ts(tC, tr, dTh + 2 * ks, id2, acc);                    // hi . hi, tf32
            mma_ts(tS, tr + 8, dTh + 2 * ks, id2, acc);
            mma_ts_bf16(tC, tr + 16, dTc + 2 * ks, id2b, 1);           // hi.lo + lo.hi, bf16, K = 16 = 8 rows x (hi, lo)
            mma_ts_bf16(tS, tr + 24, dTc + 2 * ks, id2b, 1);
            // The h = 0 warps get their half of R[g] (rows 0..15) back after 8 of the 16 MMAs: the wait for MMA2(b - 2)
            // was 13 % of the stall samples of the SIMT warps (ncu); H pass -5 % (1.37 -> 1.29 ms at 65 536 x 32 768).
            // Handing the halves OVER separately as well (MMA2 starting on the first half alone) gave it back: 1.37 ms.
            if (ks == 1) commit(&bar_rfree[g][0]);
          }
          commit(&bar_rfree[g][1]);
          commit(&bar_empty[s]);
          if (b + 2 >= nb || ((ob + 1) % kFlush) == 0) commit(&bar_cd[g]);   // a chain ends here
        }
        __syncwarp();
        TC_EV(0, b, 3);
      }
    }
  } else {
    // ------------------------------------------------------------- SIMT warps: lane = column j, 16 rows i per block
    const int q = warp & 3, w4 = warp >> 2, g = w4 & 1, h = w4 >> 1;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int tl = q * 32 + lane;                                      // TMEM lane == column inside the CTA tile
    const int64_t col = jb + tl;
    const float eps = a.eps;
    {  // resident A operand: this thread's column of H, k = 8 w4 .. 8 w4 + 7: tf32 hi plane (32 columns) and
       // the bf16 correction plane [hi (k = 0..31) | lo (k = 0..31)], two elements per column (32 columns)
      uint32_t hi[8], bh[4], bl[4];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float* __restrict__ Hb = NBMF_TSH(a.H);
        const float x0 = Hb[(size_t)(8 * w4 + e) * a.ldh + col], x1 = Hb[(size_t)(8 * w4 + e + 1) * a.ldh + col];
        const float h0 = tf32_trunc(x0), h1 = tf32_trunc(x1);
        hi[e] = __float_as_uint(h0);
        hi[e + 1] = __float_as_uint(h1);
        bh[e >> 1] = bf16_bits(h0) | (bf16_bits(h1) << 16);
        bl[e >> 1] = bf16_bits(x0 - h0) | (bf16_bits(x1 - h1) << 16);
      }
      tmem_st8(tA + lane_off + 8 * w4, hi);
      tmem_st4(tA + 32 + lane_off + 4 * w4, bh);
      tmem_st4(tA + 48 + lane_off + 4 * w4, bl);
      wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_a);
    }
    uint32_t fmask = 0u;                                               // all ones: this column accumulates the zeros' plane
    if constexpr (FLIP) fmask = NBMF_TSH(a.flipcol)[col];
    // fp32 sums of this thread's accumulator slice (k = HALF h .. of the group's Q and S, HALF = KB / 2) live in
    // shared memory: they are touched once per chain, registers are what the hot loop is short of
    constexpr int HALF = KB / 2;
    float* __restrict__ myacc = sAcc + tid;
    if constexpr (CD) {
#pragma unroll
      for (int e = 0; e < 2 * HALF; ++e) myacc[e * 512] = 0.f;
    }
    int flushed = 0;                                                   // chains of this group already flushed
    auto flush = [&]() {                                               // TMEM chain -> fp32 accumulators
      mbar_wait(&bar_cd[g], flushed & 1);
      fence_after_sync();
#pragma unroll
      for (int part = 0; part < 2; ++part) {                           // Q, then S
        if constexpr (KB == 32) {
          uint32_t c[16];
          tmem_ld16(tAcc + 64 * g + 32 * part + lane_off + 16 * h, c);
          wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) myacc[(16 * part + e) * 512] += __uint_as_float(c[e]);
        } else {
          uint32_t c[8];
          tmem_ld8(tAcc + 64 * g + 32 * part + lane_off + 8 * h, c);
          wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) myacc[(8 * part + e) * 512] += __uint_as_float(c[e]);
        }
      }
      ++flushed;
    };
    const uint32_t* __restrict__ pc = NBMF_TSH(a.Pc) + ((size_t)blockIdx.x * a.nrb + (size_t)(r0 >> 5)) * 128 + tl;
    uint32_t word = g < nb ? pc[(size_t)g * 128] : 0u;
    const uint32_t* __restrict__ mc = nullptr;
    uint32_t mword = 0u;
    if constexpr (STRICT) {
      mc = NBMF_TSH(a.Mc) + ((size_t)blockIdx.x * a.nrb + (size_t)(r0 >> 5)) * 128 + tl;
      mword = g < nb ? mc[(size_t)g * 128] : 0u;
    }
    float ll = 0.f, ll_sum = 0.f, ll_c = 0.f;                          // fp64 is slow here: compensated fp32 sum
    // A probe of an mbarrier costs ~200 clk even when its phase completed long ago, so every wait of the
    // loop is probed early and only checked where it is needed: the latency hides behind the arithmetic.
    bool ok_theta = false;
    for (int b = g; b < nb; b += 2) {
      const int ob = b >> 1;
      const uint32_t bits = word >> (16 * h), obits = mword >> (16 * h);
      if (b + 2 < nb) {                                                // prefetch the next block's bits
        word = pc[(size_t)(b + 2) * 128];
        if constexpr (STRICT) mword = mc[(size_t)(b + 2) * 128];
      }
      uint32_t qbits = bits;                                           // the plane that is accumulated directly
      if constexpr (FLIP) {
        if constexpr (STRICT) { if (fmask) qbits = obits & ~bits; } else { qbits = bits ^ fmask; }
      }
      TC_EV(1 + g, b, 0);
      if (!ok_theta) mbar_wait(&bar_theta[g], ob & 1);
      TC_EV(1 + g, b, 1);
      fence_after_sync();
      uint32_t v[16];
      const bool ok_rfree = !(cd && b >= 2);
      {
        uint32_t v0[8], v1[8];
        tmem_ld8(tTheta + 32 * g + lane_off + 16 * h, v0);
        tmem_ld8(tTheta + 32 * g + lane_off + 16 * h + 8, v1);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[e] = v0[e]; v[8 + e] = v1[e]; }
      }
      TC_EV(1 + g, b, 2);
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tfree[g]);                       // Theta[g] may be overwritten by MMA1(b+2)
      float llb = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t out[32];
        float prod = 1.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {                                  // out: Rq_hi | R_hi | Rq_c | R_c, 8 columns each
          float x;
          if constexpr (!CD)
            x = STRICT ? h_x_strict(__uint_as_float(v[8 * u + e]), bits, obits, 1u << (8 * u + e), eps)
                       : h_x(__uint_as_float(v[8 * u + e]), bits, 1u << (8 * u + e), eps);
          else if constexpr (STRICT)
            h_entry_strict(__uint_as_float(v[8 * u + e]), bits, obits, qbits, 1u << (8 * u + e), eps, x, out[8 + e], out[24 + e], out[e], out[16 + e]);
          else if constexpr (FLIP)
            h_entry_q(__uint_as_float(v[8 * u + e]), bits, qbits, 1u << (8 * u + e), eps, x, out[8 + e], out[24 + e], out[e], out[16 + e]);
          else
            h_entry(__uint_as_float(v[8 * u + e]), bits, 1u << (8 * u + e), eps, x, out[8 + e], out[24 + e], out[e], out[16 + e]);
          prod = (e & 3) ? prod * x : x;
          if ((e & 3) == 3) llb += logu_(prod);                        // eps >= 1e-9: four factors cannot underflow
        }
        if (cd) {
          if (u == 0) {                                                // MMA2(b-2) must be done reading R[g]
            TC_EV(1 + g, b, 3);
            if (!ok_rfree) mbar_wait(&bar_rfree[g][h], (ob - 1) & 1);
            fence_after_sync();
            TC_EV(1 + g, b, 4);
          }
          tmem_st32(tR + 128 * g + lane_off + 32 * (2 * h + u), out);
        }
      }
      ok_theta = b + 2 < nb && mbar_try(smem_u32(&bar_theta[g]), (ob + 1) & 1);
      TC_EV(1 + g, b, 5);
      if (cd) {
        if (ob > 0 && (ob % kFlush) == 0) flush();                     // previous chain: its MMAs ended a block ago
        wait_st();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s[g]);
      TC_EV(1 + g, b, 6);
      ll += llb;
      if ((ob & 15) == 15) {                                           // Kahan step every 16 own blocks
        const float y = ll - ll_c, t = ll_sum + y;
        ll_c = (t - ll_sum) - y;
        ll_sum = t;
        ll = 0.f;
      }
    }
    ll_total = ((double)ll_sum - (double)ll_c) + (double)ll;
    if (col >= a.n) ll_total = 0.0;
    if (cd) {
      if (g < nb) flush();                                             // the group's last chain
      asm volatile("bar.sync 1, 512;" ::: "memory");                   // the 16 SIMT warps only
      if (g == 0) {                                                    // group 0 + group 1 (thread tid + 128), fixed order
        float* __restrict__ base = NBMF_TSH(a.CD) + (size_t)(split * 2) * 32 * a.ldh;   // C rows 0..31 then D rows 0..31
#pragma unroll
        for (int e = 0; e < HALF; ++e) {                               // column j of row k: coalesced across the warp
          const float qv = myacc[e * 512] + myacc[e * 512 + 128];
          const float sm = myacc[(HALF + e) * 512] + myacc[(HALF + e) * 512 + 128];
          const float other = sm - qv;
          base[(size_t)(HALF * h + e) * a.ldh + col] = (FLIP && fmask) ? other : qv;
          base[(size_t)(32 + HALF * h + e) * a.ldh + col] = (FLIP && fmask) ? qv : other;
        }
        if constexpr (KB == 16) {                                      // rows 16..31 of the K <= 32 layout: padding
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            base[(size_t)(16 + 8 * h + e) * a.ldh + col] = 0.f;
            base[(size_t)(48 + 8 * h + e) * a.ldh + col] = 0.f;
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == TC_MMA1_WARP) tmem_dealloc(tb, 512);
  const double tot = block_sum<TC_THREADS>(ll_total, red_scratch);
  if (tid == 0) NBMF_TSH(a.LL)[(size_t)split * gridDim.x + blockIdx.x] = tot * log_unit<float>();
}

// =====================================================================================
// W pass (TMEM lane = row i).  CTA = 128 rows i, streams 64-column blocks of H.
//   MMA1: Theta'[128 i x 64 j] = W[128 x 32 k] . Ht[64 j x 32 k]^T
//   SIMT: signed ratio s = 1/(+-x) on observed entries (= p - q), q-sum of the zeros' 1/x, s -> TMEM (hi, lo)
//   MMA2: G[128 i x 32 k] += S[128 x 64 j] . H[32 k x 64 j]^T
// TMEM: A hi 0..31, bf16 [hi|lo] 32..63 | Theta[g] 64..191 | S[g] 192..447 (per 8 columns: S_hi S_c) | G[g] 448..511
// =====================================================================================
constexpr int WTC_STAGES = 6;
constexpr int WTC_LOOK = 1;
constexpr int WTC_STAGE_BYTES = 32768;
constexpr int WTC_OFF_X = WTC_STAGES * WTC_STAGE_BYTES;           // group 1 -> group 0 accumulator exchange [16][256]
constexpr int WTC_OFF_Q = WTC_OFF_X + 16 * 256 * 4;
constexpr int WTC_SMEM = WTC_OFF_Q + 4 * 128 * 4 + 1024;

template <int KB>
__global__ void __launch_bounds__(TC_THREADS, 1) w_pass_tc_kernel(const WTcArgs a) {
  static_assert(KB == 16 || KB == 32, "KB");
  using namespace tc;
  if (*NBMF_TSH(a.done)) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sX = reinterpret_cast<float*>(smem + WTC_OFF_X);           // [16 accumulators][256 threads of group 1]
  float* sQ = reinterpret_cast<float*>(smem + WTC_OFF_Q);           // [4 (group, half)][128 lanes]
  __shared__ uint64_t bar_full[WTC_STAGES], bar_empty[WTC_STAGES];
  __shared__ uint64_t bar_a, bar_theta[2], bar_tfree[2], bar_s[2], bar_sfree[2][2], bar_g[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t ib = (int64_t)blockIdx.x * 128;
  const int64_t c0 = (int64_t)blockIdx.y * a.cols_per_split;
  const int64_t c1 = min(a.n, c0 + a.cols_per_split);
  const int nb = c1 > c0 ? (int)((c1 - c0 + 63) / 64) : 0;

  if (warp == TC_MMA1_WARP) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < WTC_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_a, TC_SIMT_WARPS);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bar_theta[g], 1); mbar_init(&bar_tfree[g], TC_SIMT_WARPS / 2);
      mbar_init(&bar_s[g], TC_SIMT_WARPS / 2); mbar_init(&bar_sfree[g][0], 1); mbar_init(&bar_sfree[g][1], 1); mbar_init(&bar_g[g], 1);
    }
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tA = tb, tTheta = tb + 64, tS = tb + 192, tG = tb + 448;
  constexpr uint32_t id1 = idesc_tf32(128, 64), id1b = idesc_bf16(128, 64), id2 = idesc_tf32(128, KB), id2b = idesc_bf16(128, KB);
  constexpr int KS = KB / 8, NG = KB / 2;                              // K steps of MMA1; G columns one SIMT thread owns

  if (warp == TC_MMA1_WARP || warp == TC_MMA1_WARP + 1) {
    // ------------------------------------------------------------- MMA1 issuer of group g + producer of its stages
    // (one 32 KB bulk copy per block, one own block ahead; see the H pass)
    const int g = warp - TC_MMA1_WARP;
    const bool leader = elect_one();
    const uint32_t tT = tTheta + 64 * g;
    const float* src = NBMF_TSH(a.Hf) + (size_t)(c0 >> 6) * 8192;
    auto produce = [&](int bp) {
      if (bp >= nb) return;
      const int s = bp % WTC_STAGES;
      if (bp >= WTC_STAGES) mbar_wait(&bar_empty[s], ((bp / WTC_STAGES) - 1) & 1);
      if (leader) {
        mbar_expect_tx(&bar_full[s], WTC_STAGE_BYTES);
        bulk_g2s(smem + s * WTC_STAGE_BYTES, src + (size_t)bp * 8192, WTC_STAGE_BYTES, &bar_full[s]);
      }
      __syncwarp();
    };
#pragma unroll
    for (int i = 0; i < WTC_LOOK; ++i) produce(g + 2 * i);
    mbar_wait(&bar_a, 0);
    fence_after_sync();
    for (int b = g; b < nb; b += 2) {
      const int s = b % WTC_STAGES;
      mbar_wait(&bar_full[s], (b / WTC_STAGES) & 1);
      if (b >= 2) mbar_wait(&bar_tfree[g], ((b >> 1) - 1) & 1);
      fence_after_sync();
      if (leader) {
        const uint32_t st = smem_u32(smem + s * WTC_STAGE_BYTES);
        const uint64_t dHh = desc_kmajor_sw128(st), dHc = desc_kmajor_sw128(st + 8192);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma_ts(tT, tA + 8 * ks, dHh + 2 * ks, id1, ks > 0);          // hi . hi, tf32
#pragma unroll
        for (int i = 0; i < KS; ++i) {                                 // hi.lo + lo.hi, bf16 (KB = 16: see the H pass)
          const int ks = KB == 16 ? 2 * i : i;
          mma_ts_bf16(tT, tA + 32 + 8 * ks, dHc + 2 * ks, id1b, 1);
        }
        commit(&bar_theta[g]);
      }
      __syncwarp();
      produce(b + 2 * WTC_LOOK);
    }
  } else if (warp == TC_MMA2_WARP || warp == TC_MMA2_WARP + 1) {
    // ------------------------------------------------------------- MMA2 issuer of group g
    const int g = warp - TC_MMA2_WARP;
    const bool leader = elect_one();
    const uint32_t tSb = tS + 128 * g, tGa = tG + 32 * g;
    for (int b = g; b < nb; b += 2) {
      const int ob = b >> 1;
      const bool chain_start = (ob % kFlush) == 0;
      mbar_wait(&bar_s[g], ob & 1);
      fence_after_sync();
      if (leader) {
        const int s = b % WTC_STAGES;
        const uint32_t st = smem_u32(smem + s * WTC_STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dBh = desc_kmajor_sw128(st + 16384 + (ks >> 2) * 4096) + 2 * (ks & 3);
          const uint64_t dBc = desc_kmajor_sw128(st + 24576 + (ks >> 2) * 4096) + 2 * (ks & 3);
          const uint32_t ts = tSb + 16 * ks;                             // S_hi +0, S_c +8
          mma_ts(tGa, ts, dBh, id2, (ks > 0 || !chain_start) ? 1u : 0u); // hi . hi, tf32
          mma_ts_bf16(tGa, ts + 8, dBc, id2b, 1);                        // hi.lo + lo.hi, bf16
          if (ks == 3) commit(&bar_sfree[g][0]);                         // columns 0..31 (the h = 0 warps' half of S[g]) are read
        }
        commit(&bar_sfree[g][1]);
        commit(&bar_empty[s]);
        if (b + 2 >= nb || ((ob + 1) % kFlush) == 0) commit(&bar_g[g]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------- SIMT warps: lane = row i, 32 columns j per block
    const int q = warp & 3, w4 = warp >> 2, g = w4 & 1, h = w4 >> 1;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int tl = q * 32 + lane;
    const int64_t row = ib + tl;
    const float eps2 = 2.0f * a.eps;                                   // Theta arrives with eps added: see w_entry
    {  // resident A operand: this thread's row of W, k = 8 w4 .. 8 w4 + 7 (zero beyond m)
      float x[8];
      if (row < a.m) {
        const float* __restrict__ Wb = NBMF_TSH(a.W);
        const float4 x0 = *reinterpret_cast<const float4*>(Wb + (size_t)row * 32 + 8 * w4);
        const float4 x1 = *reinterpret_cast<const float4*>(Wb + (size_t)row * 32 + 8 * w4 + 4);
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
      tmem_st8(tA + lane_off + 8 * w4, hi);
      tmem_st4(tA + 32 + lane_off + 4 * w4, bh);
      tmem_st4(tA + 48 + lane_off + 4 * w4, bl);
      wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_a);
    }
    float accG[NG];                                                    // k = NG h .. NG h + NG - 1 of the group's G, fp32
#pragma unroll
    for (int e = 0; e < NG; ++e) accG[e] = 0.f;
    int flushed = 0;
    auto flush = [&]() {                                               // TMEM chain -> fp32 register accumulators
      mbar_wait(&bar_g[g], flushed & 1);
      fence_after_sync();
      if constexpr (KB == 32) {
        uint32_t c[16];
        tmem_ld16(tG + 32 * g + lane_off + 16 * h, c);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) accG[e] += __uint_as_float(c[e]);
      } else {
        uint32_t c[8];
        tmem_ld8(tG + 32 * g + lane_off + 8 * h, c);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) accG[e] += __uint_as_float(c[e]);
      }
      ++flushed;
    };
    const uint2* __restrict__ pm = NBMF_TSH(a.PM) + ((size_t)blockIdx.x * a.wpr + (size_t)(c0 >> 5) + h) * 128 + tl;
    uint2 word = g < nb ? pm[(size_t)(2 * g) * 128] : make_uint2(0u, 0u);
    float qsum = 0.f;
    bool ok_theta = false;                                             // early barrier probes, see the H pass
    for (int b = g; b < nb; b += 2) {
      const int ob = b >> 1;
      const uint2 bits = word;
      if (b + 2 < nb) word = pm[(size_t)(2 * (b + 2)) * 128];
      if (!ok_theta) mbar_wait(&bar_theta[g], ob & 1);
      fence_after_sync();
      float qb = 0.f;                                                  // this block's sum over the observed zeros of 1/x
      bool ok_sfree = true;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t v[16];
        tmem_ld16(tTheta + 64 * g + lane_off + 32 * h + 16 * u, v);
        if (u == 0 && b >= 2) ok_sfree = mbar_try(smem_u32(&bar_sfree[g][h]), (ob - 1) & 1);
        wait_ld();
        if (u == 1) {                                                  // Theta[g] may be overwritten by MMA1(b+2)
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_tfree[g]);
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {                                  // 8 entries = one K step of MMA2: S_hi | S_lo
          uint32_t out[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float sv;
            w_entry(__uint_as_float(v[8 * w + e]), bits.x, bits.y, 1u << (16 * u + 8 * w + e), eps2, sv, qb);
            const float hi = tf32_trunc(sv);
            out[e] = __float_as_uint(hi);
            out[8 + e] = pack_bf16x2(hi, sv - hi);                       // (hi, lo) of s in one column
          }
          if (u == 0 && w == 0) {                                      // MMA2(b-2) must be done reading S[g]
            if (!ok_sfree) mbar_wait(&bar_sfree[g][h], (ob - 1) & 1);
            fence_after_sync();
          }
          tmem_st16(tS + 128 * g + lane_off + 16 * (4 * h + 2 * u + w), out);
        }
      }
      ok_theta = b + 2 < nb && mbar_try(smem_u32(&bar_theta[g]), (ob + 1) & 1);
      if (ob > 0 && (ob % kFlush) == 0) flush();                       // previous chain: its MMAs ended a block ago
      wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s[g]);
      qsum += qb;
    }
    if (g < nb) flush();
    sQ[w4 * 128 + tl] = qsum;
    const int t = h * 128 + tl;                                        // same (lane, k range) in both groups
    if (g == 1) {
#pragma unroll
      for (int e = 0; e < NG; ++e) sX[e * 256 + t] = accG[e];
    }
    asm volatile("bar.sync 1, 512;" ::: "memory");                     // the 16 SIMT warps only
    if (g == 0 && row < a.m) {                                         // group 0 + group 1, fixed order
      float* __restrict__ Gg = NBMF_TSH(a.G) + ((size_t)blockIdx.y * a.m + row) * 32 + NG * h;
#pragma unroll
      for (int e = 0; e < NG; e += 4)
        *reinterpret_cast<float4*>(Gg + e) =
            make_float4(accG[e] + sX[e * 256 + t], accG[e + 1] + sX[(e + 1) * 256 + t],
                        accG[e + 2] + sX[(e + 2) * 256 + t], accG[e + 3] + sX[(e + 3) * 256 + t]);
      if (h == 0)
        NBMF_TSH(a.Q)[(size_t)blockIdx.y * a.m + row] = ((sQ[tl] + sQ[128 + tl]) + sQ[256 + tl]) + sQ[384 + tl];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == TC_MMA1_WARP) tmem_dealloc(tb, 512);
}

template <int KB>
inline void launch_w_pass_tc(const WTcArgs& a, int nsplit, int batch_n, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set{0};
  ensure_dynamic_smem(w_pass_tc_kernel<KB>, WTC_SMEM, attr_set);
  dim3 grid((unsigned)((a.m + 127) / 128), (unsigned)nsplit, (unsigned)batch_n);
  w_pass_tc_kernel<KB><<<grid, TC_THREADS, WTC_SMEM, st>>>(a);
}
// Decides, per H pass, which columns accumulate the zeros' plane directly (density of ones above the mean of the column
// of H, see h_pass_tc_kernel) and whether any does.  One thread per column (K coalesced loads); the grid-wide "any" is an
// integer OR collected by the last block to finish (order-independent: deterministic).  flag[0] = result, flag[1] =
// OR in progress, flag[2] = blocks finished; the last block resets [1], [2] for the next launch.
__global__ void __launch_bounds__(256) flip_cols_kernel(const float* __restrict__ H, int64_t ldh, int64_t n, int k, int64_t m,
                                                        const uint32_t* __restrict__ colcnt, uint32_t* __restrict__ flipcol,
                                                        int* __restrict__ flag, const int* __restrict__ done, int64_t bstride) {
  if (bstride) {                                  // fit blockIdx.y of a batch: its own workspace
    const size_t sh = (size_t)blockIdx.y * (size_t)bstride;
    H = batch_shift(H, sh); colcnt = batch_shift(colcnt, sh); flipcol = batch_shift(flipcol, sh);
    flag = batch_shift(flag, sh); done = batch_shift(done, sh);
  }
  if (*done) return;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t f = 0u;
  if (j < n) {
    float hsum = 0.f;
    for (int kk = 0; kk < k; ++kk) hsum += H[(size_t)kk * ldh + j];
    if ((float)colcnt[j] * (float)k > hsum * (float)m) f = 0xffffffffu;        // density > mean H
  }
  if (j < ldh) flipcol[j] = f;
  const int any = __syncthreads_or(f != 0u);
  if (threadIdx.x == 0) {
    if (any) atomicOr(&flag[1], 1);
    __threadfence();
    if (atomicAdd(&flag[2], 1) == (int)gridDim.x - 1) {                        // last block: publish and reset
      __threadfence();
      flag[0] = atomicExch(&flag[1], 0);
      flag[2] = 0;
    }
  }
}

template <int KB>
inline void launch_h_pass_tc(const HTcArgs& a, int nsplit, int batch_n, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set[6];
  ensure_dynamic_smem(h_pass_tc_kernel<KB, false, true, false>, HTC_SMEM, attr_set[0]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, true, true, false>, HTC_SMEM, attr_set[1]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, false, false, false>, HTC_SMEM, attr_set[2]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, true, false, false>, HTC_SMEM, attr_set[3]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, false, true, true>, HTC_SMEM, attr_set[4]);
  ensure_dynamic_smem(h_pass_tc_kernel<KB, true, true, true>, HTC_SMEM, attr_set[5]);
  dim3 grid((unsigned)((a.n + 127) / 128), (unsigned)nsplit, (unsigned)batch_n);
  if (a.compute_cd) {
    const bool flips = a.colcnt != nullptr && a.flipcol != nullptr && a.flip_any != nullptr;
    HTcArgs b = a;
    if (!flips) b.flip_any = nullptr;                                  // no decision data: the FLIP = false kernel always runs
    else flip_cols_kernel<<<dim3((unsigned)((a.ldh + 255) / 256), (unsigned)batch_n), 256, 0, st>>>(
        a.H, a.ldh, a.n, a.k, a.m, a.colcnt, const_cast<uint32_t*>(a.flipcol), const_cast<int*>(a.flip_any), a.done, a.batch_stride);
    if (a.Mc) h_pass_tc_kernel<KB, true, true, false><<<grid, TC_THREADS, HTC_SMEM, st>>>(b);
    else h_pass_tc_kernel<KB, false, true, false><<<grid, TC_THREADS, HTC_SMEM, st>>>(b);
    if (flips) {
      if (a.Mc) h_pass_tc_kernel<KB, true, true, true><<<grid, TC_THREADS, HTC_SMEM, st>>>(b);
      else h_pass_tc_kernel<KB, false, true, true><<<grid, TC_THREADS, HTC_SMEM, st>>>(b);
    }
  } else {
    HTcArgs b = a;
    b.flip_any = nullptr;
    if (a.Mc) h_pass_tc_kernel<KB, true, false, false><<<grid, TC_THREADS, HTC_SMEM, st>>>(b);
    else h_pass_tc_kernel<KB, false, false, false><<<grid, TC_THREADS, HTC_SMEM, st>>>(b);
  }
}

}  // namespace k32
}  // namespace nbmf

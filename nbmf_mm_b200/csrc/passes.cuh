// The two fused pass kernels of one MM iteration (reference: src/nbmf_mm/_solver.py:5-59
// for the update, :148-162 for the loss that rides on the H pass).
//
// Both are structured like attention with fp32/fp64 SIMT math: the Theta = W.H tile is
// formed on the fly from a factor tile staged in shared memory and a factor slice held
// in registers, the masked ratios are computed in registers, and the second contraction
// is accumulated in registers.  Theta, the ratios and the weights never touch HBM.
//
// Register tiling.  A warp is split into G = 32/S "groups" of S lanes.  A group owns C
// output vectors (columns j for the H pass, rows i for the W pass); inside a group the K
// axis is split S ways (lane kp owns k in [kp*KH, (kp+1)*KH), KH = KP/S).  Theta partial
// dot products are combined with log2(S) xor-shuffles.  Accumulators are (k, k+1) pairs so
// that fp32 issues packed FFMA2 (one issue slot per two FMAs).
#pragma once
#include <cuda_fp16.h>

#include <type_traits>
#include "args.h"
#include "common.cuh"

namespace nbmf {

template <typename Real, int KH>
__device__ __forceinline__ void load_pairs(const Real* __restrict__ src, typename Vec2<Real>::type (&dst)[KH / 2]) {
  using V2 = typename Vec2<Real>::type;
  if constexpr (sizeof(Real) == 4 && (KH % 4) == 0) {
#pragma unroll
    for (int q = 0; q < KH / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(src + 4 * q);
      dst[2 * q] = make2(v.x, v.y);
      dst[2 * q + 1] = make2(v.z, v.w);
    }
  } else {
#pragma unroll
    for (int q = 0; q < KH / 2; ++q) dst[q] = *reinterpret_cast<const V2*>(src + 2 * q);
  }
}

// N consecutive elements of the dense V*mask layout (storage type VT: Real or __half), 16 bytes per load where the
// run is long enough; p is aligned to the run (column offsets are multiples of N, rows of 1024 elements).
template <typename VT, typename Real, int N>
__device__ __forceinline__ void load_v(const VT* __restrict__ p, Real (&out)[N]) {
  constexpr int B = (int)sizeof(VT) * N;
  VT tmp[N];
  if constexpr (B % 16 == 0) {
#pragma unroll
    for (int i = 0; i < B / 16; ++i) reinterpret_cast<uint4*>(tmp)[i] = reinterpret_cast<const uint4*>(p)[i];
  } else if constexpr (B == 8) {
    *reinterpret_cast<uint2*>(tmp) = *reinterpret_cast<const uint2*>(p);
  } else if constexpr (B == 4) {
    *reinterpret_cast<uint32_t*>(tmp) = *reinterpret_cast<const uint32_t*>(p);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = p[i];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = (Real)tmp[i];
}

// Factor rows in shared memory: lane kp of a group reads its K slice (KH elements) of a row with 16-byte loads, and the
// 8 lanes of one shared-memory phase hold min(S, 8) different slices.  Slices that lie a multiple of 128 bytes apart fall
// on the same banks (ncu: 5e8 bank conflicts per pass in fp64, where two slices of 16 doubles are exactly 128 bytes apart;
// 8-way at K = 128), so such tilings get 16 bytes of padding after every slice: slice kp starts at kp * (slice + 16).
// MINS: pad only tilings with at least this many slices per row.  The W pass pads every conflicting tiling (fp64 K <= 32,
// two slices: 75.5 -> 71.6 ms), the H pass from four slices on (fp32 K = 64: 181.6 -> 174.8 ms): its two-slice fp64 case
// is not bound by shared memory and lost 3 % to the longer address arithmetic.
template <typename Real, int KH, int S, int MINS>
struct SlicePad {
  static constexpr int SLICE_B = KH * (int)sizeof(Real);
  static constexpr int ways() {
    int w = 0;
    for (int d = 0; d < (S < 8 ? S : 8); ++d)
      if ((d * SLICE_B) % 128 == 0) ++w;
    return w;
  }
  static constexpr int PAD_B = (ways() >= 2 && S >= MINS) ? 16 : 0;
  static constexpr int STRIDE_B = SLICE_B + PAD_B;          // bytes from one slice of a row to the next
  static constexpr int ROW_B = S * STRIDE_B;                // bytes per factor row in shared memory
  static constexpr int CPR = S * SLICE_B / 16;              // 16-byte chunks per row of the contiguous global layout
  static_assert((S * SLICE_B) % 16 == 0, "factor rows must be 16-byte multiples");
  static_assert(PAD_B == 0 || SLICE_B % 16 == 0, "padded slices must be 16-byte multiples");
  // byte offset inside the tile of chunk c of the contiguous global layout (row-major, K elements per row)
  __device__ static __forceinline__ int chunk_offset(int c) {
    if constexpr (PAD_B == 0) return 16 * c;
    constexpr int CH = SLICE_B / 16;                        // chunks per slice
    const int r = c / CPR, q = c - r * CPR;
    return r * ROW_B + q * 16 + (q / CH) * PAD_B;
  }
};

// =====================================================================================
// H pass:  C[k][j] = sum_i W[i][k] * pos[i][j] / (Theta[i][j] + eps)
//          D[k][j] = sum_i W[i][k] * neg[i][j] / ((1 - Theta[i][j]) + eps)
//          LL      = sum_ij pos*log(Theta+eps) + neg*log((1-Theta)+eps)      (_solver.py:39-43,150-154)
// pos = V&mask; neg = 1 - pos (reference quirk) or mask - pos (STRICT).
// A thread owns C contiguous columns x KH values of k; rows stream through shared memory.
// =====================================================================================
template <typename Real_, int KP_, int S_, int C_, int NW_, int MINB_, bool DENSE_, bool STRICT_, typename VT_ = Real_>
struct HCfg {
  using Real = Real_;
  using VT = VT_;                            // storage type of dense V*mask: Real, or __half (fp16 layout)
  static constexpr int KP = KP_, S = S_, C = C_, NW = NW_, MINB = MINB_;
  static constexpr bool DENSE = DENSE_, STRICT = STRICT_;
  static constexpr int KH = KP / S;
  static constexpr int G = 32 / S;
  static constexpr int CW = G * C;          // columns per warp
  static constexpr int BN = NW * CW;        // columns per CTA
  static constexpr int WPT = BN / 32;       // bit words per tile row
  static constexpr int BM = DENSE ? 16 : 32;   // rows per stage (dense: the V tile rides in the stage too)
  static constexpr int NSTAGE = 3;
  static constexpr int NT = NW * 32;
  using Pad = SlicePad<Real, KP / S, S, 4>;
  static constexpr int W_BYTES = BM * Pad::ROW_B;
  static constexpr int P_BYTES = DENSE ? 0 : BM * WPT * 4;
  static constexpr int M_BYTES = STRICT ? BM * WPT * 4 : 0;
  static constexpr int V_BYTES = DENSE ? BM * BN * (int)sizeof(VT) : 0;   // dense V*mask tile, staged by cp.async
  static constexpr int STAGE_BYTES = W_BYTES + P_BYTES + M_BYTES + V_BYTES;
  static constexpr int SMEM = NSTAGE * STAGE_BYTES;
  static_assert(KP % S == 0 && KH % 2 == 0, "K slice per lane must be even");
  static_assert(BN % 64 == 0 && 1024 % BN == 0, "column tile must divide the 1024-column pitch");
  static constexpr int BIT_CB = WPT >= 4 ? 16 : 8;   // bytes per cp.async of a bit-plane tile row (BN = 64: two words)
  static_assert(32 % C == 0, "a thread's bits must not straddle a word");
};

template <typename Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) h_pass_kernel(const HPassArgs a) {
  using Real = typename Cfg::Real;
  using V2 = typename Vec2<Real>::type;
  constexpr int KP = Cfg::KP, S = Cfg::S, C = Cfg::C, KH = Cfg::KH, NT = Cfg::NT;
  constexpr int BM = Cfg::BM, WPT = Cfg::WPT, NSTAGE = Cfg::NSTAGE;
  constexpr bool DENSE = Cfg::DENSE, STRICT = Cfg::STRICT;
  // fit blockIdx.z of a batch works in its own workspace: W, H, CD, LL and done are shifted where they are used (the
  // offset is recomputed there: it must not occupy registers across the main loop); 0 for a single fit
#define NBMF_BSH ((size_t)blockIdx.z * (size_t)a.batch_stride)
  if (*batch_shift(a.done, NBMF_BSH)) return;

  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_scratch[NT / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / S, kp = lane % S;
  const int jw = warp * Cfg::CW + g * C;                        // column offset inside the CTA tile
  const int64_t j0 = (int64_t)blockIdx.x * Cfg::BN + jw;         // first owned column
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * a.rows_per_split;
  const int64_t r1 = min(a.m, r0 + a.rows_per_split);
  const Real eps = (Real)a.eps;

  // ---- H slice -> registers (padded columns hold 0.5, always in bounds)
  V2 hp[C][KH / 2];
  {
    const Real* __restrict__ Hg = batch_shift(reinterpret_cast<const Real*>(a.H), NBMF_BSH);
#pragma unroll
    for (int q = 0; q < KH / 2; ++q)
#pragma unroll
      for (int cc = 0; cc < C; ++cc)
        hp[cc][q] = make2(Hg[(int64_t)(kp * KH + 2 * q) * a.ldh + j0 + cc],
                          Hg[(int64_t)(kp * KH + 2 * q + 1) * a.ldh + j0 + cc]);
  }
  V2 cacc[C][KH / 2], dacc[C][KH / 2];
#pragma unroll
  for (int cc = 0; cc < C; ++cc)
#pragma unroll
    for (int q = 0; q < KH / 2; ++q) cacc[cc][q] = dacc[cc][q] = make2(Real(0), Real(0));
  double lld[C];
#pragma unroll
  for (int cc = 0; cc < C; ++cc) lld[cc] = 0.0;

  const unsigned char* __restrict__ Wg = batch_shift(reinterpret_cast<const unsigned char*>(a.W), NBMF_BSH);
  using VT = typename Cfg::VT;
  const VT* __restrict__ Vg = reinterpret_cast<const VT*>(a.Vm);
  const int64_t ntiles = (r1 > r0) ? (r1 - r0 + BM - 1) / BM : 0;

  auto issue_tile = [&](int64_t t) {
    unsigned char* st = smem + (size_t)(t % NSTAGE) * Cfg::STAGE_BYTES;
    const int64_t rb = r0 + t * BM;
    const int nrows = (int)min((int64_t)BM, r1 - rb);
    const int wchunks = nrows * (KP * (int)sizeof(Real) / 16);
    const unsigned char* wsrc = Wg + (size_t)rb * KP * sizeof(Real);
    for (int c = tid; c < wchunks; c += NT) cp_async16(st + Cfg::Pad::chunk_offset(c), wsrc + 16 * (size_t)c);
    auto bit_rows = [&](unsigned char* dst, const uint32_t* __restrict__ plane) {   // nrows x WPT words of a bit plane
      constexpr int CB = Cfg::BIT_CB, CPR = WPT * 4 / CB;         // chunks of CB bytes per tile row
      const int bchunks = nrows * CPR;
      for (int c = tid; c < bchunks; c += NT) {
        const int r = c / CPR, q = c % CPR;
        const uint32_t* src = plane + (size_t)(rb + r) * a.wpr + (size_t)blockIdx.x * WPT + (CB / 4) * q;
        if constexpr (CB == 16) cp_async16(dst + CB * c, src);
        else cp_async8(dst + CB * c, src);
      }
    };
    if constexpr (!DENSE) bit_rows(st + Cfg::W_BYTES, a.P);
    if constexpr (STRICT) bit_rows(st + Cfg::W_BYTES + Cfg::P_BYTES, a.M);
    if constexpr (DENSE) {
      // the V*mask tile: rows of BN elements, NSTAGE - 1 tiles (2 x 16 rows) ahead of the arithmetic -- a register
      // prefetch one row ahead left the DRAM latency exposed (52 % of the stall samples on the load's first use)
      constexpr int CPR = Cfg::BN * (int)sizeof(VT) / 16;         // 16-byte chunks per tile row
      const int vchunks = nrows * CPR;
      const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(Vg + (size_t)blockIdx.x * Cfg::BN);
      for (int c = tid; c < vchunks; c += NT) {
        const int r = c / CPR, q = c % CPR;
        cp_async16(st + Cfg::W_BYTES + Cfg::P_BYTES + Cfg::M_BYTES + 16 * c,
                   vsrc + ((size_t)(rb + r) * a.ldv) * sizeof(VT) + 16 * (size_t)q);
      }
    }
  };

#pragma unroll
  for (int t = 0; t < NSTAGE - 1; ++t) {
    if (t < ntiles) issue_tile(t);
    cp_async_commit();
  }

  // the tile loop, once per value of a.compute_cd: as a run-time test inside the row loop the flag is a (uniform) branch
  // that cuts the loop body into basic blocks, and ptxas then cannot interleave the Theta dots, the reciprocal chain and
  // the accumulations of the two unrolled rows
  auto tile_loop = [&](auto cd_tag) {
  constexpr bool CD = decltype(cd_tag)::value;
  for (int64_t t = 0; t < ntiles; ++t) {
    if (t + NSTAGE - 1 < ntiles) issue_tile(t + NSTAGE - 1);
    cp_async_commit();
    cp_async_wait<NSTAGE - 1>();
    __syncthreads();

    const unsigned char* st = smem + (size_t)(t % NSTAGE) * Cfg::STAGE_BYTES;
    const unsigned char* Wt = st + kp * Cfg::Pad::STRIDE_B;      // this lane's K slice of row 0
    const uint32_t* Pb = reinterpret_cast<const uint32_t*>(st + Cfg::W_BYTES);
    const uint32_t* Mb = reinterpret_cast<const uint32_t*>(st + Cfg::W_BYTES + Cfg::P_BYTES);
    const int64_t rb = r0 + t * BM;
    const int nrows = (int)min((int64_t)BM, r1 - rb);

    Real ll[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) ll[cc] = Real(0);
    // fp64, binary V: log() is ~100 instructions, as much as the rest of an entry.  The log of a product of LOGG
    // factors replaces LOGG logs (x >= eps, so 8 factors stay far above the fp64 underflow threshold; the product's
    // rounding error is ~1e-15 relative).  Negative factors are flagged so that a negative x still gives NaN, as
    // the reference's log(x) would, instead of cancelling against another negative factor.
    constexpr bool LOGPROD = !DENSE && sizeof(Real) == 8;
    constexpr int LOGG = 8;
    Real px[LOGPROD ? C : 1];
    uint32_t negm = 0u;                                   // bit cc: a factor of column cc was negative
    if constexpr (LOGPROD) {
#pragma unroll
      for (int cc = 0; cc < C; ++cc) px[cc] = Real(1);
    }
    auto flush_logs = [&]() {
      if constexpr (LOGPROD) {
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
          ll[cc] += logu_(((negm >> cc) & 1u) ? Real(NAN) : px[cc]);
          px[cc] = Real(1);
        }
        negm = 0u;
      }
    };
    const VT* Vt = reinterpret_cast<const VT*>(st + Cfg::W_BYTES + Cfg::P_BYTES + Cfg::M_BYTES) + jw;   // dense V tile

#pragma unroll 2
    for (int r = 0; r < nrows; ++r) {
      V2 wp[KH / 2];
      load_pairs<Real, KH>(reinterpret_cast<const Real*>(Wt + r * Cfg::Pad::ROW_B), wp);

      // ---- Theta[i][j] for the C owned columns
      V2 th[C];
#pragma unroll
      for (int cc = 0; cc < C; ++cc) th[cc] = make2(Real(0), Real(0));
#pragma unroll
      for (int q = 0; q < KH / 2; ++q)
#pragma unroll
        for (int cc = 0; cc < C; ++cc) th[cc] = fma2(wp[q], hp[cc][q], th[cc]);
      Real theta[C];
#pragma unroll
      for (int cc = 0; cc < C; ++cc) {
        theta[cc] = th[cc].x + th[cc].y;
#pragma unroll
        for (int o = 1; o < S; o <<= 1) theta[cc] += __shfl_xor_sync(0xffffffffu, theta[cc], o);
      }

      uint32_t pb = 0, mb = 0xffffffffu;
      if constexpr (!DENSE) pb = Pb[r * WPT + (jw >> 5)] >> (jw & 31);
      if constexpr (STRICT) mb = Mb[r * WPT + (jw >> 5)] >> (jw & 31);
      Real vcur[C];
      if constexpr (DENSE) load_v<VT, Real, C>(Vt + r * Cfg::BN, vcur);

      // ---- masked ratios in registers, loss term, second contraction
#pragma unroll
      for (int cc = 0; cc < C; ++cc) {
        Real rp, rn_;
        if constexpr (!DENSE) {
          const bool p = (pb >> cc) & 1u;
          const Real x = (p ? theta[cc] : (Real(1) - theta[cc])) + eps;
          Real r_ = rcp_(x);
          if constexpr (LOGPROD) {
            Real xf = x;
            if constexpr (STRICT) {
              const bool o = (mb >> cc) & 1u;
              r_ = o ? r_ : Real(0);
              xf = o ? x : Real(1);
            }
            px[cc] *= xf;
            negm |= (xf < Real(0)) ? (1u << cc) : 0u;
          } else {
            Real lg = logu_(x);
            if constexpr (STRICT) {
              const bool o = (mb >> cc) & 1u;
              r_ = o ? r_ : Real(0);
              lg = o ? lg : Real(0);
            }
            ll[cc] += lg;
          }
          rp = p ? r_ : Real(0);
          rn_ = p ? Real(0) : r_;
        } else {
          const Real v = vcur[cc];
          const Real xp = theta[cc] + eps;
          const Real xn = (Real(1) - theta[cc]) + eps;
          Real neg;
          if constexpr (STRICT) neg = (((mb >> cc) & 1u) ? Real(1) : Real(0)) - v;
          else neg = Real(1) - v;
          rp = div_(v, xp);
          rn_ = div_(neg, xn);
          ll[cc] += v * logu_(xp) + neg * logu_(xn);
        }
        if constexpr (CD) {
          const V2 rp2 = make2(rp, rp), rn2 = make2(rn_, rn_);
#pragma unroll
          for (int q = 0; q < KH / 2; ++q) {
            cacc[cc][q] = fma2(wp[q], rp2, cacc[cc][q]);
            dacc[cc][q] = fma2(wp[q], rn2, dacc[cc][q]);
          }
        }
      }
      if constexpr (LOGPROD) {
        if ((r & (LOGG - 1)) == LOGG - 1) flush_logs();
      }
    }
    if constexpr (LOGPROD) {
      if (nrows & (LOGG - 1)) flush_logs();
    }
#pragma unroll
    for (int cc = 0; cc < C; ++cc) lld[cc] += (double)ll[cc];
    __syncthreads();
  }
  };
  if (a.compute_cd) tile_loop(std::true_type{});
  else tile_loop(std::false_type{});
  cp_async_wait<0>();

  // ---- partial C, D for this row split
  if (a.compute_cd) {
    Real* __restrict__ Cg = batch_shift(reinterpret_cast<Real*>(a.CD), NBMF_BSH) + (size_t)(split * 2 + 0) * KP * a.ldh;
    Real* __restrict__ Dg = batch_shift(reinterpret_cast<Real*>(a.CD), NBMF_BSH) + (size_t)(split * 2 + 1) * KP * a.ldh;
#pragma unroll
    for (int q = 0; q < KH / 2; ++q) {
      const size_t o0 = (size_t)(kp * KH + 2 * q) * a.ldh + j0;
      const size_t o1 = o0 + a.ldh;
#pragma unroll
      for (int cc = 0; cc < C; ++cc) {
        Cg[o0 + cc] = cacc[cc][q].x;
        Cg[o1 + cc] = cacc[cc][q].y;
        Dg[o0 + cc] = dacc[cc][q].x;
        Dg[o1 + cc] = dacc[cc][q].y;
      }
    }
  }
  // ---- partial log-likelihood: one lane per group counts, padded columns excluded
  double mine = 0.0;
  if (kp == 0) {
#pragma unroll
    for (int cc = 0; cc < C; ++cc)
      if (j0 + cc < a.n) mine += lld[cc];
  }
  const double tot = block_sum<NT>(mine, red_scratch);
  if (tid == 0) batch_shift(a.LL, NBMF_BSH)[(size_t)split * gridDim.x + blockIdx.x] = tot * log_unit<Real>();
}

template <typename Cfg>
void launch_h_pass(const HPassArgs& a, int nsplit, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set{0};
  ensure_dynamic_smem(h_pass_kernel<Cfg>, Cfg::SMEM, attr_set);
  dim3 grid((unsigned)((a.n + Cfg::BN - 1) / Cfg::BN), (unsigned)nsplit, (unsigned)a.batch_n);
  h_pass_kernel<Cfg><<<grid, Cfg::NT, Cfg::SMEM, st>>>(a);
}

// =====================================================================================
// W pass:  Gs[i][k] = sum_j H[k][j] * (p_ij - q_ij),   Qs[i] = sum_j q_ij
//   p = posW/(Theta+eps), q = negW/((1-Theta)+eps)  with posW = V&mask, negW = mask&~V
//   (_solver.py:50-53, always properly masked), so that
//   G[i][k] = sum_j H[k][j] p + (1-H[k][j]) q = Gs[i][k] + Qs[i].
// For binary V exactly one of p, q is non-zero, so p - q is exact.
// A thread owns C rows x KH values of k; columns stream through shared memory (Ht tiles).
// =====================================================================================
template <typename Real_, int KP_, int S_, int C_, int NW_, int MINB_, bool DENSE_, typename VT_ = Real_>
struct WCfg {
  using Real = Real_;
  using VT = VT_;
  static constexpr int KP = KP_, S = S_, C = C_, NW = NW_, MINB = MINB_;
  static constexpr bool DENSE = DENSE_;
  static constexpr int KH = KP / S;
  static constexpr int G = 32 / S;
  static constexpr int RW = G * C;          // rows per warp
  static constexpr int BMR = NW * RW;       // rows per CTA
  static constexpr int BNT = 128;           // columns per stage
  static constexpr int NWORD = BNT / 32;
  static constexpr int NT = NW * 32;
  using Pad = SlicePad<Real, KP / S, S, 2>;
  static constexpr int HT_BYTES = BNT * Pad::ROW_B;
  static constexpr int P_BYTES = DENSE ? 0 : BMR * NWORD * 4;
  static constexpr int M_BYTES = BMR * NWORD * 4;
  static constexpr int STAGE_BYTES = HT_BYTES + P_BYTES + M_BYTES;
  // two stages (the next Ht tile loads while this one is worked on) unless they exceed the SM's shared memory: fp64 at
  // K > 96 (128 x 128 x 8 = 128 KB per tile) runs single-staged
  static constexpr int NSTAGE = 2 * STAGE_BYTES <= 200 * 1024 ? 2 : 1;
  static constexpr int SMEM = NSTAGE * STAGE_BYTES;
  static_assert(KP % S == 0 && KH % 2 == 0, "K slice per lane must be even");
};

template <typename Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) w_pass_kernel(const WPassArgs a) {
  using Real = typename Cfg::Real;
  using V2 = typename Vec2<Real>::type;
  constexpr int KP = Cfg::KP, S = Cfg::S, C = Cfg::C, KH = Cfg::KH, NT = Cfg::NT;
  constexpr int BNT = Cfg::BNT, NWORD = Cfg::NWORD, NSTAGE = Cfg::NSTAGE, BMR = Cfg::BMR;
  constexpr bool DENSE = Cfg::DENSE;
  if (*batch_shift(a.done, NBMF_BSH)) return;                    // batch: W, Ht, G, Q, done shifted at their uses

  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / S, kp = lane % S;
  const int il = warp * Cfg::RW + g * C;                          // first owned row inside the CTA tile
  const int64_t ib = (int64_t)blockIdx.x * BMR;
  const int split = blockIdx.y;
  const int64_t c0 = (int64_t)split * a.cols_per_split;
  const int64_t c1 = min(a.n, c0 + a.cols_per_split);
  const bool has_mask = (a.M != nullptr);
  const Real eps = (Real)a.eps;

  // ---- W slice -> registers (rows past m are clamped; their results are not stored)
  V2 wp[C][KH / 2];
  {
    const Real* __restrict__ Wg = batch_shift(reinterpret_cast<const Real*>(a.W), NBMF_BSH);
#pragma unroll
    for (int rr = 0; rr < C; ++rr) {
      const int64_t row = min(ib + il + rr, a.m - 1);
      load_pairs<Real, KH>(Wg + (size_t)row * KP + kp * KH, wp[rr]);
    }
  }
  V2 gacc[C][KH / 2];
  Real qacc[C];
#pragma unroll
  for (int rr = 0; rr < C; ++rr) {
    qacc[rr] = Real(0);
#pragma unroll
    for (int q = 0; q < KH / 2; ++q) gacc[rr][q] = make2(Real(0), Real(0));
  }

  const unsigned char* __restrict__ Htg = batch_shift(reinterpret_cast<const unsigned char*>(a.Ht), NBMF_BSH);
  using VT = typename Cfg::VT;
  const VT* __restrict__ Vg = reinterpret_cast<const VT*>(a.Vm);
  // weighted (non-0/1) observation mask: its VALUES, dense, same layout as V*mask (the reference multiplies by them:
  // (1 - Y).T * mask.T, _solver.py:32); NULL for a 0/1 mask, whose bit plane says it all
  const Real* __restrict__ Wmg = DENSE ? reinterpret_cast<const Real*>(a.Wm) : nullptr;
  const int64_t ntiles = (c1 > c0) ? (c1 - c0 + BNT - 1) / BNT : 0;

  auto issue_tile = [&](int64_t t) {
    unsigned char* st = smem + (size_t)(t % NSTAGE) * Cfg::STAGE_BYTES;
    const int64_t cb = c0 + t * BNT;
    constexpr int HCH = BNT * Cfg::Pad::CPR;
    const unsigned char* hsrc = Htg + (size_t)cb * KP * sizeof(Real);
    for (int c = tid; c < HCH; c += NT) cp_async16(st + Cfg::Pad::chunk_offset(c), hsrc + 16 * (size_t)c);
    const int64_t wb = cb >> 5;                                   // first bit word of this tile
    for (int r = tid; r < BMR; r += NT) {
      const int64_t row = min(ib + r, a.m - 1);
      if constexpr (!DENSE) cp_async16(st + Cfg::HT_BYTES + 16 * r, a.P + (size_t)row * a.wpr + wb);
      if (has_mask) cp_async16(st + Cfg::HT_BYTES + Cfg::P_BYTES + 16 * r, a.M + (size_t)row * a.wpr + wb);
    }
  };

  constexpr bool VPREF = DENSE && sizeof(Real) == 4;              // fp64 variants are short of registers: no lookahead
  Real vnext[VPREF ? C : 1][8];                                   // dense V: the chunk after the one being worked on
  if constexpr (VPREF) {
#pragma unroll
    for (int rr = 0; rr < C; ++rr)
      load_v<VT, Real, 8>(Vg + (size_t)min(ib + il + rr, a.m - 1) * a.ldv + min(c0, a.ldv - 8), vnext[rr]);
  }
  if constexpr (NSTAGE > 1) {
    if (ntiles > 0) issue_tile(0);
    cp_async_commit();
  }

  for (int64_t t = 0; t < ntiles; ++t) {
    if constexpr (NSTAGE > 1) {
      if (t + 1 < ntiles) issue_tile(t + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {                                                       // single stage: load, then work (see WCfg)
      issue_tile(t);
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();

    const unsigned char* st = smem + (size_t)(t % NSTAGE) * Cfg::STAGE_BYTES;
    const unsigned char* Hts = st + kp * Cfg::Pad::STRIDE_B;    // this lane's K slice of column 0 of the tile
    const uint32_t* Pb = reinterpret_cast<const uint32_t*>(st + Cfg::HT_BYTES);
    const uint32_t* Mb = reinterpret_cast<const uint32_t*>(st + Cfg::HT_BYTES + Cfg::P_BYTES);
    const int64_t cb = c0 + t * BNT;

#pragma unroll 1
    for (int w = 0; w < NWORD; ++w) {
      const int64_t colw = cb + 32 * w;
      if (colw >= c1) break;
      // columns past n (or past this split) contribute nothing
      const int64_t rem = c1 - colw;
      const uint32_t valid = rem >= 32 ? 0xffffffffu : ((1u << (int)rem) - 1u);
      uint32_t pw[C], mw[C];
#pragma unroll
      for (int rr = 0; rr < C; ++rr) {
        pw[rr] = DENSE ? 0u : Pb[(il + rr) * NWORD + w];
        mw[rr] = (has_mask ? Mb[(il + rr) * NWORD + w] : 0xffffffffu) & valid;
      }
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        Real vv[DENSE ? C : 1][8];
        if constexpr (VPREF) {
          // this chunk was loaded while the previous one was being worked on; start the next one now (the chunks
          // of a column split are consecutive runs of 8 columns, whatever tile / word they fall in)
#pragma unroll
          for (int rr = 0; rr < C; ++rr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) vv[rr][e] = vnext[rr][e];
            const int64_t row = min(ib + il + rr, a.m - 1);
            load_v<VT, Real, 8>(Vg + (size_t)row * a.ldv + min(colw + 8 * u + 8, a.ldv - 8), vnext[rr]);
          }
        } else if constexpr (DENSE) {
#pragma unroll
          for (int rr = 0; rr < C; ++rr)
            load_v<VT, Real, 8>(Vg + (size_t)min(ib + il + rr, a.m - 1) * a.ldv + colw + 8 * u, vv[rr]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int jj = 8 * u + e;
          V2 hq[KH / 2];
          load_pairs<Real, KH>(reinterpret_cast<const Real*>(Hts + (32 * w + jj) * Cfg::Pad::ROW_B), hq);
          V2 th[C];
#pragma unroll
          for (int rr = 0; rr < C; ++rr) th[rr] = make2(Real(0), Real(0));
#pragma unroll
          for (int q = 0; q < KH / 2; ++q)
#pragma unroll
            for (int rr = 0; rr < C; ++rr) th[rr] = fma2(wp[rr][q], hq[q], th[rr]);
#pragma unroll
          for (int rr = 0; rr < C; ++rr) {
            Real theta = th[rr].x + th[rr].y;
#pragma unroll
            for (int o = 1; o < S; o <<= 1) theta += __shfl_xor_sync(0xffffffffu, theta, o);
            const bool ob = (mw[rr] >> jj) & 1u;
            Real s;
            if constexpr (!DENSE) {
              const bool p = (pw[rr] >> jj) & 1u;
              const Real x = (p ? theta : (Real(1) - theta)) + eps;
              Real r_ = rcp_(x);
              r_ = ob ? r_ : Real(0);
              s = p ? r_ : -r_;
              qacc[rr] += p ? Real(0) : r_;
            } else {
              const Real v = vv[rr][e];
              Real wv = ob ? Real(1) : Real(0);
              if (Wmg) wv = Wmg[(size_t)min(ib + il + rr, a.m - 1) * a.ldv + colw + jj];
              const Real pa = div_(v, theta + eps);
              const Real qb = div_(wv - v, (Real(1) - theta) + eps);     // (1 - V) * mask = mask - V * mask
              s = pa - qb;
              qacc[rr] += qb;
            }
            const V2 s2 = make2(s, s);
#pragma unroll
            for (int q = 0; q < KH / 2; ++q) gacc[rr][q] = fma2(hq[q], s2, gacc[rr][q]);
          }
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  Real* __restrict__ Gg = batch_shift(reinterpret_cast<Real*>(a.G), NBMF_BSH) + (size_t)split * a.m * KP;
  Real* __restrict__ Qg = batch_shift(reinterpret_cast<Real*>(a.Q), NBMF_BSH) + (size_t)split * a.m;
#pragma unroll
  for (int rr = 0; rr < C; ++rr) {
    const int64_t row = ib + il + rr;
    if (row < a.m) {
#pragma unroll
      for (int q = 0; q < KH / 2; ++q)
        *reinterpret_cast<V2*>(Gg + (size_t)row * KP + kp * KH + 2 * q) = gacc[rr][q];
      if (kp == 0) Qg[row] = qacc[rr];
    }
  }
}

template <typename Cfg>
void launch_w_pass(const WPassArgs& a, int nsplit, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set{0};
  ensure_dynamic_smem(w_pass_kernel<Cfg>, Cfg::SMEM, attr_set);
  dim3 grid((unsigned)((a.m + Cfg::BMR - 1) / Cfg::BMR), (unsigned)nsplit, (unsigned)a.batch_n);
  w_pass_kernel<Cfg><<<grid, Cfg::NT, Cfg::SMEM, st>>>(a);
}

}  // namespace nbmf

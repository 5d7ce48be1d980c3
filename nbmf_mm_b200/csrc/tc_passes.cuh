// Tensor-core (tcgen05 / TMEM) versions of the two pass kernels for K <= 32, fp32, bit-packed V.
//
// Same math as passes.cuh, restructured like attention on Blackwell (S = QK^T -> P -> O = PV):
//   MMA1  Theta tile = factor tile . streamed factor block          (tcgen05.mma, SS, accumulate in TMEM)
//   SIMT  tcgen05.ld Theta -> masked ratio in registers -> tcgen05.st back to TMEM
//   MMA2  accumulator (+)= ratio[TMEM] . streamed factor block       (tcgen05.mma, TS: A operand from TMEM)
// fp32 accuracy comes from a 3-term TF32 split: x = hi + lo with hi = tf32(x) (the tensor core reads
// only the top 19 bits), lo = x - hi, and a.b ~ hi.hi + hi.lo + lo.hi (measured ~1.5e-6 relative,
// tools/tc_probe.cu).  Operand blocks are pre-split and pre-swizzled in global memory by the epilogue
// kernels (format_factors.cu), so every shared-memory stage is filled by plain 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx): no tensor maps, no swizzling in the hot loop.
//
// One CTA = 4 SIMT warps (thread t owns TMEM lane t) + 1 control warp (one elected thread issues the
// bulk copies and the MMAs).  TMEM use is <= 256 columns so two CTAs share an SM and overlap each
// other's MMA and SIMT phases.
#pragma once
#include "args.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace nbmf {

struct TcFactors {
  const float* Ha;   // [ldh/64][hi|lo][64 rows j x 32 k]            Ht blocks, K-major, SW128
  const float* Hb;   // [ldh/64][hi|lo][2 K-blocks][32 rows k x 32 j] H blocks
  const float* Wa;   // [mpad/64][hi|lo][64 rows i x 32 k]           W blocks
  const float* Wb;   // [mpad/64][hi|lo][2 K-blocks][32 rows k x 32 i] W^T blocks
};

struct WTcArgs {
  TcFactors f;
  const uint32_t* P;
  const uint32_t* M;
  int64_t m, n, wpr;
  int64_t cols_per_split;    // multiple of 64
  float* G;                  // [nsplit][m][32]
  float* Q;                  // [nsplit][m]
  float eps;
  const int* done;
};

struct HTcArgs {
  TcFactors f;
  const uint32_t* Pt;        // transposed plane: [n][wpr_t], bit i of row j = P[i][j]
  int64_t m, n, ldh, wpr_t;
  int64_t rows_per_split;    // multiple of 32
  float* CD;                 // [nsplit][2][32][ldh]
  double* LL;                // [nsplit * gridDim.x]
  float eps;
  const int* done;
  int compute_cd;
};

constexpr int TC_THREADS = 160;          // 4 SIMT warps + 1 control warp
constexpr int TC_BLK_FLOATS = 64 * 32;   // one hi (or lo) 64-row block

// =====================================================================================
// W pass.  CTA = 128 rows i.  Per 64-column block:
//   MMA1: Theta'[128 i x 64 j] = W[128x32] . Ht[64x32]^T
//   SIMT: s = +-1/x on observed entries (p - q, exact), q-sum in registers, s -> TMEM (hi, lo)
//   MMA2: G[128 i x 32 k] += S[128 x 64 j] . H[32 k x 64 j]^T
// TMEM columns: Theta 0..63 | S_hi 64..127 | S_lo 128..191 | G 192..223
// =====================================================================================
constexpr int WTC_STAGE_BYTES = 32768;                       // Ha hi|lo (16 KB) + Hb hi|lo (16 KB)
constexpr int WTC_SMEM = 32768 + 2 * WTC_STAGE_BYTES + 1024; // W tile hi|lo + 2 stages + alignment slack

__global__ void __launch_bounds__(TC_THREADS, 2) w_pass_tc_kernel(const WTcArgs a) {
  using namespace tc;
  if (*a.done) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sW = smem;                                  // [hi 16 KB][lo 16 KB], 128 rows each
  __shared__ uint64_t bar_w, bar_full[2], bar_empty[2], bar_theta, bar_s, bar_g;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t ib = (int64_t)blockIdx.x * 128;
  const int64_t c0 = (int64_t)blockIdx.y * a.cols_per_split;
  const int64_t c1 = min(a.n, c0 + a.cols_per_split);
  const int nb = (int)((c1 - c0 + 63) / 64);

  if (warp == 4) tmem_alloc(&tmem_base_s, 256);
  if (tid == 128) {
    mbar_init(&bar_w, 1); mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
    mbar_init(&bar_theta, 1); mbar_init(&bar_s, 128); mbar_init(&bar_g, 1);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tTheta = tb, tS = tb + 64, tSl = tb + 128, tG = tb + 192;

  if (tid == 128) {
    // ------------------------------------------------------------- control thread
    const int64_t wblk = ib / 64;
    mbar_expect_tx(&bar_w, 32768);
    bulk_g2s(sW, a.f.Wa + (size_t)(wblk * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_w);                   // rows 0..63 hi
    bulk_g2s(sW + 8192, a.f.Wa + (size_t)((wblk + 1) * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_w);      // rows 64..127 hi
    bulk_g2s(sW + 16384, a.f.Wa + (size_t)(wblk * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_w);           // lo
    bulk_g2s(sW + 24576, a.f.Wa + (size_t)((wblk + 1) * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_w);
    auto produce = [&](int b) {
      const int s = b & 1;
      if (b >= 2) mbar_wait(&bar_empty[s], ((b >> 1) - 1) & 1);
      unsigned char* st = smem + 32768 + s * WTC_STAGE_BYTES;
      const int64_t hblk = c0 / 64 + b;
      mbar_expect_tx(&bar_full[s], 32768);
      bulk_g2s(st, a.f.Ha + (size_t)hblk * 2 * TC_BLK_FLOATS, 16384, &bar_full[s]);
      bulk_g2s(st + 16384, a.f.Hb + (size_t)hblk * 2 * TC_BLK_FLOATS, 16384, &bar_full[s]);
    };
    const uint64_t dWh = desc_kmajor_sw128(smem_u32(sW)), dWl = desc_kmajor_sw128(smem_u32(sW + 16384));
    constexpr uint32_t id1 = idesc_tf32(128, 64), id2 = idesc_tf32(128, 32);
    auto mma1 = [&](int b) {
      const int s = b & 1;
      mbar_wait(&bar_full[s], (b >> 1) & 1);
      fence_after_sync();
      unsigned char* st = smem + 32768 + s * WTC_STAGE_BYTES;
      const uint64_t dHh = desc_kmajor_sw128(smem_u32(st)), dHl = desc_kmajor_sw128(smem_u32(st + 8192));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tTheta, dWh + 2 * ks, dHh + 2 * ks, id1, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tTheta, dWh + 2 * ks, dHl + 2 * ks, id1, 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tTheta, dWl + 2 * ks, dHh + 2 * ks, id1, 1);
      commit(&bar_theta);
    };
    if (nb > 0) produce(0);
    if (nb > 1) produce(1);
    mbar_wait(&bar_w, 0);
    if (nb > 0) mma1(0);
    for (int b = 0; b < nb; ++b) {
      const int s = b & 1;
      unsigned char* st = smem + 32768 + s * WTC_STAGE_BYTES;
      mbar_wait(&bar_s, b & 1);
      fence_after_sync();
#pragma unroll
      for (int t = 0; t < 3; ++t) {                            // S.H, S.H_lo, S_lo.H
        const uint32_t ta = (t == 2) ? tSl : tS;
        unsigned char* hb = st + 16384 + (t == 1 ? 8192 : 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dB = desc_kmajor_sw128(smem_u32(hb + (ks >> 2) * 4096)) + 2 * (ks & 3);
          mma_ts(tG, ta + 8 * ks, dB, id2, (b | t | ks) > 0);
        }
      }
      commit(&bar_empty[s]);
      if (b + 1 < nb) mma1(b + 1);
      if (b + 2 < nb) produce(b + 2);
    }
    commit(&bar_g);
  } else if (tid < 128) {
    // ------------------------------------------------------------- SIMT threads: one matrix row each
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int64_t row = ib + tid;
    const int64_t rowc = min(row, a.m - 1);
    const uint32_t* __restrict__ Prow = a.P + (size_t)rowc * a.wpr;
    const uint32_t* __restrict__ Mrow = a.M ? a.M + (size_t)rowc * a.wpr : nullptr;
    const float eps = a.eps;
    float qacc = 0.f;
    for (int b = 0; b < nb; ++b) {
      const int64_t colb = c0 + 64 * (int64_t)b;
      const int64_t wi = colb >> 5;
      uint32_t pw[2], mw[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t rem = c1 - (colb + 32 * h);
        const uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
        pw[h] = Prow[wi + h];
        mw[h] = (Mrow ? Mrow[wi + h] : 0xffffffffu) & valid;
      }
      mbar_wait(&bar_theta, b & 1);
      fence_after_sync();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[16], sh[16], sl[16];
        tmem_ld16(tTheta + lane_off + 16 * c, v);
        wait_ld();
        const uint32_t pbits = pw[c >> 1] >> ((c & 1) * 16), mbits = mw[c >> 1] >> ((c & 1) * 16);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float theta = __uint_as_float(v[e]);
          const bool p = (pbits >> e) & 1u, o = (mbits >> e) & 1u;
          const float x = (p ? theta : (1.0f - theta)) + eps;
          float r = rcp_(x);
          r = o ? r : 0.0f;
          const float s = p ? r : -r;
          qacc += p ? 0.0f : r;
          const float hi = tc::tf32_trunc(s);
          sh[e] = __float_as_uint(hi);
          sl[e] = __float_as_uint(s - hi);
        }
        tmem_st16(tS + lane_off + 16 * c, sh);
        tmem_st16(tSl + lane_off + 16 * c, sl);
      }
      wait_st();
      fence_before_sync();
      mbar_arrive(&bar_s);
    }
    mbar_wait(&bar_g, 0);
    fence_after_sync();
    uint32_t g0[16], g1[16];                                   // warp-collective loads: all 32 lanes take part
    tmem_ld16(tG + lane_off, g0);
    tmem_ld16(tG + lane_off + 16, g1);
    wait_ld();
    if (row < a.m) {
      float* __restrict__ Gg = a.G + ((size_t)blockIdx.y * a.m + row) * 32;
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        *reinterpret_cast<float4*>(Gg + e) =
            make_float4(__uint_as_float(g0[e]), __uint_as_float(g0[e + 1]), __uint_as_float(g0[e + 2]), __uint_as_float(g0[e + 3]));
        *reinterpret_cast<float4*>(Gg + 16 + e) =
            make_float4(__uint_as_float(g1[e]), __uint_as_float(g1[e + 1]), __uint_as_float(g1[e + 2]), __uint_as_float(g1[e + 3]));
      }
      a.Q[(size_t)blockIdx.y * a.m + row] = qacc;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, 256);
}

// =====================================================================================
// H pass (transposed ownership: thread t owns column j, TMEM lane t).  CTA = 128 columns j.
// Per 32-row block:
//   MMA1: Theta^T[128 j x 32 i] = Ht[128x32] . W[32x32]^T
//   SIMT: bits from the transposed plane Pt; rp / rn ratios (hi, lo) -> TMEM; fused NLL in registers
//   MMA2: C^T[128 j x 32 k] += Rp[128 x 32 i] . W^T[32 k x 32 i]^T   (and D^T with Rn)
// TMEM columns: Theta 0..31 | Rp_hi 32..63 | Rp_lo 64..95 | Rn_hi 96..127 | Rn_lo 128..159 | C 160..191 | D 192..223
// =====================================================================================
constexpr int HTC_STAGE_BYTES = 16384;                       // Wa 32 rows hi|lo (8 KB) + Wb K-block hi|lo (8 KB)
constexpr int HTC_SMEM = 32768 + 2 * HTC_STAGE_BYTES + 1024;

__global__ void __launch_bounds__(TC_THREADS, 2) h_pass_tc_kernel(const HTcArgs a) {
  using namespace tc;
  if (*a.done) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sH = smem;                                  // Ht tile [hi 16 KB][lo 16 KB], 128 rows j
  __shared__ uint64_t bar_h, bar_full[2], bar_empty[2], bar_theta, bar_s, bar_g;
  __shared__ uint32_t tmem_base_s;
  __shared__ double red_scratch[TC_THREADS / 32];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t jb = (int64_t)blockIdx.x * 128;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * a.rows_per_split;
  const int64_t r1 = min(a.m, r0 + a.rows_per_split);
  const int nb = (int)((r1 - r0 + 31) / 32);
  const bool cd = a.compute_cd != 0;

  if (warp == 4) tmem_alloc(&tmem_base_s, 256);
  if (tid == 128) {
    mbar_init(&bar_h, 1); mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
    mbar_init(&bar_theta, 1); mbar_init(&bar_s, 128); mbar_init(&bar_g, 1);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tTheta = tb, tRp = tb + 32, tRpl = tb + 64, tRn = tb + 96, tRnl = tb + 128, tC = tb + 160, tD = tb + 192;

  double ll_total = 0.0;
  if (tid == 128) {
    // ------------------------------------------------------------- control thread
    const int64_t hblk = jb / 64;
    mbar_expect_tx(&bar_h, 32768);
    bulk_g2s(sH, a.f.Ha + (size_t)(hblk * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_h);
    bulk_g2s(sH + 8192, a.f.Ha + (size_t)((hblk + 1) * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_h);
    bulk_g2s(sH + 16384, a.f.Ha + (size_t)(hblk * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_h);
    bulk_g2s(sH + 24576, a.f.Ha + (size_t)((hblk + 1) * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_h);
    auto produce = [&](int b) {
      const int s = b & 1;
      if (b >= 2) mbar_wait(&bar_empty[s], ((b >> 1) - 1) & 1);
      unsigned char* st = smem + 32768 + s * HTC_STAGE_BYTES;
      const int64_t row = r0 + 32 * (int64_t)b;
      const int64_t wblk = row / 64;
      const int sub = (int)((row % 64) / 32);                  // which 32-row half of the 64-row block
      const float* wa = a.f.Wa + (size_t)wblk * 2 * TC_BLK_FLOATS;
      const float* wb = a.f.Wb + (size_t)wblk * 2 * TC_BLK_FLOATS;
      mbar_expect_tx(&bar_full[s], 16384);
      bulk_g2s(st, wa + sub * 1024, 4096, &bar_full[s]);                            // W rows hi
      bulk_g2s(st + 4096, wa + TC_BLK_FLOATS + sub * 1024, 4096, &bar_full[s]);     // W rows lo
      bulk_g2s(st + 8192, wb + sub * 1024, 4096, &bar_full[s]);                     // W^T K-block hi
      bulk_g2s(st + 12288, wb + TC_BLK_FLOATS + sub * 1024, 4096, &bar_full[s]);    // W^T K-block lo
    };
    const uint64_t dHh = desc_kmajor_sw128(smem_u32(sH)), dHl = desc_kmajor_sw128(smem_u32(sH + 16384));
    constexpr uint32_t id = idesc_tf32(128, 32);
    auto mma1 = [&](int b) {
      const int s = b & 1;
      mbar_wait(&bar_full[s], (b >> 1) & 1);
      fence_after_sync();
      unsigned char* st = smem + 32768 + s * HTC_STAGE_BYTES;
      const uint64_t dWh = desc_kmajor_sw128(smem_u32(st)), dWl = desc_kmajor_sw128(smem_u32(st + 4096));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tTheta, dHh + 2 * ks, dWh + 2 * ks, id, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tTheta, dHh + 2 * ks, dWl + 2 * ks, id, 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tTheta, dHl + 2 * ks, dWh + 2 * ks, id, 1);
      commit(&bar_theta);
    };
    if (nb > 0) produce(0);
    if (nb > 1) produce(1);
    mbar_wait(&bar_h, 0);
    if (nb > 0) mma1(0);
    for (int b = 0; b < nb; ++b) {
      const int s = b & 1;
      unsigned char* st = smem + 32768 + s * HTC_STAGE_BYTES;
      mbar_wait(&bar_s, b & 1);
      fence_after_sync();
      if (cd) {
        const uint64_t dTh = desc_kmajor_sw128(smem_u32(st + 8192)), dTl = desc_kmajor_sw128(smem_u32(st + 12288));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (b | ks) > 0;
          mma_ts(tC, tRp + 8 * ks, dTh + 2 * ks, id, acc);
          mma_ts(tD, tRn + 8 * ks, dTh + 2 * ks, id, acc);
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(tC, tRp + 8 * ks, dTl + 2 * ks, id, 1);
          mma_ts(tD, tRn + 8 * ks, dTl + 2 * ks, id, 1);
          mma_ts(tC, tRpl + 8 * ks, dTh + 2 * ks, id, 1);
          mma_ts(tD, tRnl + 8 * ks, dTh + 2 * ks, id, 1);
        }
        commit(&bar_empty[s]);
      } else {
        mbar_arrive(&bar_empty[s]);                            // loss-only pass: nothing reads the stage after MMA1
      }
      if (b + 1 < nb) mma1(b + 1);
      if (b + 2 < nb) produce(b + 2);
    }
    commit(&bar_g);
  } else if (tid < 128) {
    // ------------------------------------------------------------- SIMT threads: one matrix column each
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int64_t col = jb + tid;
    const int64_t colc = min(col, a.n - 1);
    const uint32_t* __restrict__ Pcol = a.Pt + (size_t)colc * a.wpr_t;
    const float eps = a.eps;
    for (int b = 0; b < nb; ++b) {
      const uint32_t pbits = Pcol[(r0 >> 5) + b];
      mbar_wait(&bar_theta, b & 1);
      fence_after_sync();
      float ll = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[16], ph[16], pl[16], nh[16], nl[16];
        tmem_ld16(tTheta + lane_off + 16 * c, v);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float theta = __uint_as_float(v[e]);
          const bool p = (pbits >> (16 * c + e)) & 1u;
          const float x = (p ? theta : (1.0f - theta)) + eps;
          const float r = rcp_(x);
          ll += logu_(x);
          const float hi = tc::tf32_trunc(r);
          const float lo = r - hi;
          ph[e] = __float_as_uint(p ? hi : 0.0f);
          pl[e] = __float_as_uint(p ? lo : 0.0f);
          nh[e] = __float_as_uint(p ? 0.0f : hi);
          nl[e] = __float_as_uint(p ? 0.0f : lo);
        }
        if (cd) {
          tmem_st16(tRp + lane_off + 16 * c, ph);
          tmem_st16(tRpl + lane_off + 16 * c, pl);
          tmem_st16(tRn + lane_off + 16 * c, nh);
          tmem_st16(tRnl + lane_off + 16 * c, nl);
        }
      }
      ll_total += (double)ll;
      if (cd) wait_st();
      fence_before_sync();
      mbar_arrive(&bar_s);
    }
    mbar_wait(&bar_g, 0);
    fence_after_sync();
    if (cd) {
      float* __restrict__ Cg = a.CD + (size_t)(split * 2 + 0) * 32 * a.ldh;
      float* __restrict__ Dg = a.CD + (size_t)(split * 2 + 1) * 32 * a.ldh;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t vc[16], vd[16];
        tmem_ld16(tC + lane_off + 16 * c, vc);
        tmem_ld16(tD + lane_off + 16 * c, vd);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) {                         // column j of row k: coalesced across the warp
          Cg[(size_t)(16 * c + e) * a.ldh + col] = __uint_as_float(vc[e]);
          Dg[(size_t)(16 * c + e) * a.ldh + col] = __uint_as_float(vd[e]);
        }
      }
    }
    if (col >= a.n) ll_total = 0.0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, 256);
  const double tot = block_sum<TC_THREADS>(ll_total, red_scratch);
  if (tid == 0) a.LL[(size_t)split * gridDim.x + blockIdx.x] = tot * log_unit<float>();
}

inline void launch_w_pass_tc(const WTcArgs& a, int nsplit, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(w_pass_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WTC_SMEM);
    attr_set = true;
  }
  dim3 grid((unsigned)((a.m + 127) / 128), (unsigned)nsplit);
  w_pass_tc_kernel<<<grid, TC_THREADS, WTC_SMEM, st>>>(a);
}
inline void launch_h_pass_tc(const HTcArgs& a, int nsplit, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(h_pass_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HTC_SMEM);
    attr_set = true;
  }
  dim3 grid((unsigned)((a.n + 127) / 128), (unsigned)nsplit);
  h_pass_tc_kernel<<<grid, TC_THREADS, HTC_SMEM, st>>>(a);
}

}  // namespace nbmf

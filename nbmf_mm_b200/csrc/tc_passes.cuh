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
// One persistent-style CTA per SM: 16 SIMT warps + 1 control warp.  SIMT warp w works on TMEM lane
// quarter (w & 3) -- the hardware restricts a warp to lanes 32*(warp % 4).. -- and on slice (w >> 2)
// of every block's columns, so four warps per scheduler hide each other's MUFU / TMEM latencies.
// Theta and the ratio regions are double buffered in TMEM: while the SIMT warps work on block b the
// tensor pipe runs MMA2(b-1) and MMA1(b+1).  The TMEM accumulators are flushed into fp32 shared-memory
// accumulators every kFlush blocks because the tensor core's accumulate truncates: an unbroken chain of
// 3e4 accumulations drifts by ~1e-4 relative (measured), a chain of a few hundred by ~1e-6.
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

constexpr int TC_SIMT_WARPS = 16;
constexpr int TC_THREADS = (TC_SIMT_WARPS + 1) * 32;   // + control warp
constexpr int TC_CTRL_TID = TC_SIMT_WARPS * 32;
constexpr int TC_BLK_FLOATS = 64 * 32;                  // one hi (or lo) 64-row operand block
constexpr int TC_STAGES = 4;
constexpr int kFlush = 16;                              // blocks per TMEM accumulation chain

// =====================================================================================
// H pass (thread = column j = TMEM lane).  CTA = 128 columns j, streams 32-row blocks of W.
//   MMA1: Theta^T[128 j x 32 i] = Ht[128x32] . W[32x32]^T
//   SIMT: bit i of the transposed plane Pt; rp / rn ratios (hi, lo) -> TMEM; fused NLL in registers
//   MMA2: C^T[128 j x 32 k] += Rp[128 x 32 i] . W^T[32 k x 32 i]^T   (and D^T with Rn)
// TMEM: Theta[2] 0..63 | R[2] = {Rp_hi, Rp_lo, Rn_hi, Rn_lo} x 32 at 64..191, 192..319 | C 320..351 | D 352..383
// =====================================================================================
constexpr int HTC_STAGE_BYTES = 16384;                  // W rows hi|lo (8 KB) + W^T K-block hi|lo (8 KB)
constexpr int HTC_OFF_STAGE = 32768;
constexpr int HTC_OFF_ACC = HTC_OFF_STAGE + TC_STAGES * HTC_STAGE_BYTES;
constexpr int HTC_SMEM = HTC_OFF_ACC + 64 * 128 * 4 + 1024;

__global__ void __launch_bounds__(TC_THREADS, 1) h_pass_tc_kernel(const HTcArgs a) {
  using namespace tc;
  if (*a.done) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sH = smem;                                         // Ht tile [hi 16 KB][lo 16 KB], 128 rows j
  float* sAcc = reinterpret_cast<float*>(smem + HTC_OFF_ACC);       // [64 accumulator columns][128 lanes]
  __shared__ uint64_t bar_h, bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_theta[2], bar_s[2], bar_cd;
  __shared__ uint32_t tmem_base_s;
  __shared__ double red_scratch[TC_THREADS / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t jb = (int64_t)blockIdx.x * 128;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * a.rows_per_split;
  const int64_t r1 = min(a.m, r0 + a.rows_per_split);
  const int nb = r1 > r0 ? (int)((r1 - r0 + 31) / 32) : 0;
  const bool cd = a.compute_cd != 0;

  for (int e = tid; e < 64 * 128; e += TC_THREADS) sAcc[e] = 0.0f;
  if (warp == TC_SIMT_WARPS) tmem_alloc(&tmem_base_s, 512);
  if (tid == TC_CTRL_TID) {
    mbar_init(&bar_h, 1);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_theta[0], 1); mbar_init(&bar_theta[1], 1);
    mbar_init(&bar_s[0], TC_SIMT_WARPS); mbar_init(&bar_s[1], TC_SIMT_WARPS);
    mbar_init(&bar_cd, 1);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tTheta = tb, tR = tb + 64, tC = tb + 320, tD = tb + 352;
  const int nflush = nb > 0 ? (nb - 1) / kFlush : 0;                // mid-pass flushes before the final one

  double ll_total = 0.0;
  if (tid == TC_CTRL_TID) {
    // ------------------------------------------------------------- control thread
    const int64_t hblk = jb / 64;
    mbar_expect_tx(&bar_h, 32768);
    bulk_g2s(sH, a.f.Ha + (size_t)(hblk * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_h);
    bulk_g2s(sH + 8192, a.f.Ha + (size_t)((hblk + 1) * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_h);
    bulk_g2s(sH + 16384, a.f.Ha + (size_t)(hblk * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_h);
    bulk_g2s(sH + 24576, a.f.Ha + (size_t)((hblk + 1) * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_h);
    auto produce = [&](int b) {
      const int s = b & (TC_STAGES - 1);
      if (b >= TC_STAGES) mbar_wait(&bar_empty[s], ((b / TC_STAGES) - 1) & 1);
      unsigned char* st = smem + HTC_OFF_STAGE + s * HTC_STAGE_BYTES;
      const int64_t row = r0 + 32 * (int64_t)b;
      const int64_t wblk = row / 64;
      const int sub = (int)((row % 64) / 32);                       // which 32-row half of the 64-row block
      const float* wa = a.f.Wa + (size_t)wblk * 2 * TC_BLK_FLOATS;
      const float* wb = a.f.Wb + (size_t)wblk * 2 * TC_BLK_FLOATS;
      mbar_expect_tx(&bar_full[s], 16384);
      bulk_g2s(st, wa + sub * 1024, 4096, &bar_full[s]);                          // W rows hi
      bulk_g2s(st + 4096, wa + TC_BLK_FLOATS + sub * 1024, 4096, &bar_full[s]);   // W rows lo
      bulk_g2s(st + 8192, wb + sub * 1024, 4096, &bar_full[s]);                   // W^T K-block hi
      bulk_g2s(st + 12288, wb + TC_BLK_FLOATS + sub * 1024, 4096, &bar_full[s]);  // W^T K-block lo
    };
    const uint64_t dHh = desc_kmajor_sw128(smem_u32(sH)), dHl = desc_kmajor_sw128(smem_u32(sH + 16384));
    constexpr uint32_t id = idesc_tf32(128, 32);
    auto mma1 = [&](int b) {
      const int s = b & (TC_STAGES - 1);
      mbar_wait(&bar_full[s], (b / TC_STAGES) & 1);
      fence_after_sync();
      unsigned char* st = smem + HTC_OFF_STAGE + s * HTC_STAGE_BYTES;
      const uint64_t dWh = desc_kmajor_sw128(smem_u32(st)), dWl = desc_kmajor_sw128(smem_u32(st + 4096));
      const uint32_t tT = tTheta + 32 * (b & 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tT, dHh + 2 * ks, dWh + 2 * ks, id, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tT, dHh + 2 * ks, dWl + 2 * ks, id, 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tT, dHl + 2 * ks, dWh + 2 * ks, id, 1);
      commit(&bar_theta[b & 1]);
    };
    for (int b = 0; b < nb && b < 3; ++b) produce(b);
    mbar_wait(&bar_h, 0);
    if (nb > 0) mma1(0);
    if (nb > 1) mma1(1);
    for (int b = 0; b < nb; ++b) {
      const int s = b & (TC_STAGES - 1);
      unsigned char* st = smem + HTC_OFF_STAGE + s * HTC_STAGE_BYTES;
      mbar_wait(&bar_s[b & 1], (b >> 1) & 1);
      fence_after_sync();
      if (cd) {
        const uint64_t dTh = desc_kmajor_sw128(smem_u32(st + 8192)), dTl = desc_kmajor_sw128(smem_u32(st + 12288));
        const uint32_t tRp = tR + 128 * (b & 1), tRpl = tRp + 32, tRn = tRp + 64, tRnl = tRp + 96;
        const bool chain_start = (b % kFlush) == 0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (ks > 0 || !chain_start) ? 1u : 0u;
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
        if (b + 1 == nb || ((b + 1) % kFlush) == 0) commit(&bar_cd);   // a chain ends here
      } else {
        mbar_arrive(&bar_empty[s]);                                    // loss-only pass: MMA1 was the last reader
      }
      if (b + 2 < nb) mma1(b + 2);
      if (b + 3 < nb) produce(b + 3);
    }
  } else if (warp < TC_SIMT_WARPS) {
    // ------------------------------------------------------------- SIMT warps: lane = column j, 8 rows i per block
    const int q = warp & 3, sub = warp >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int tl = q * 32 + lane;                                      // TMEM lane == column inside the CTA tile
    const int64_t col = jb + tl;
    const int64_t colc = min(col, a.n - 1);
    const uint32_t* __restrict__ Pcol = a.Pt + (size_t)colc * a.wpr_t + (r0 >> 5);
    const float eps = a.eps;
    float ll = 0.f;
    auto flush = [&](int idx) {                                        // TMEM chain -> fp32 shared accumulators
      mbar_wait(&bar_cd, idx & 1);
      fence_after_sync();
      uint32_t v[16];
      tmem_ld16(tC + lane_off + 16 * sub, v);
      wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e) sAcc[(16 * sub + e) * 128 + tl] += __uint_as_float(v[e]);
    };
    for (int b = 0; b < nb; ++b) {
      const uint32_t pbits = Pcol[b] >> (8 * sub);
      if (cd && b > 0 && (b % kFlush) == 0) flush(b / kFlush - 1);
      mbar_wait(&bar_theta[b & 1], (b >> 1) & 1);
      fence_after_sync();
      uint32_t v[8], ph[8], pl[8], nh[8], nl[8];
      tmem_ld8(tTheta + 32 * (b & 1) + lane_off + 8 * sub, v);
      wait_ld();
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float theta = __uint_as_float(v[e]);
        const bool p = (pbits >> e) & 1u;
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
        const uint32_t tRp = tR + 128 * (b & 1) + lane_off + 8 * sub;
        tmem_st8(tRp, ph);
        tmem_st8(tRp + 32, pl);
        tmem_st8(tRp + 64, nh);
        tmem_st8(tRp + 96, nl);
        wait_st();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s[b & 1]);
      if ((b & 15) == 15) { ll_total += (double)ll; ll = 0.f; }
    }
    ll_total += (double)ll;
    if (cd) {
      if (nb > 0) flush(nflush);
      float* __restrict__ base = a.CD + (size_t)(split * 2) * 32 * a.ldh;   // C rows 0..31 then D rows 0..31
#pragma unroll
      for (int e = 0; e < 16; ++e)                                     // column j of row k: coalesced across the warp
        base[(size_t)(16 * sub + e) * a.ldh + col] = sAcc[(16 * sub + e) * 128 + tl];
    }
    if (col >= a.n) ll_total = 0.0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == TC_SIMT_WARPS) tmem_dealloc(tb, 512);
  const double tot = block_sum<TC_THREADS>(ll_total, red_scratch);
  if (tid == 0) a.LL[(size_t)split * gridDim.x + blockIdx.x] = tot * log_unit<float>();
}

// =====================================================================================
// W pass (thread = row i = TMEM lane).  CTA = 128 rows i, streams 64-column blocks of H.
//   MMA1: Theta'[128 i x 64 j] = W[128x32] . Ht[64x32]^T
//   SIMT: s = +-1/x on observed entries (p - q, exact), q-sum in registers, s -> TMEM (hi, lo)
//   MMA2: G[128 i x 32 k] += S[128 x 64 j] . H[32 k x 64 j]^T
// TMEM: Theta[2] 0..127 | S[2] = {S_hi, S_lo} x 64 at 128..255, 256..383 | G 384..415
// =====================================================================================
constexpr int WTC_STAGE_BYTES = 32768;                  // Ha hi|lo (16 KB) + Hb hi|lo (16 KB)
constexpr int WTC_OFF_STAGE = 32768;
constexpr int WTC_OFF_ACC = WTC_OFF_STAGE + TC_STAGES * WTC_STAGE_BYTES;
constexpr int WTC_OFF_Q = WTC_OFF_ACC + 32 * 128 * 4;
constexpr int WTC_SMEM = WTC_OFF_Q + 4 * 128 * 4 + 1024;

__global__ void __launch_bounds__(TC_THREADS, 1) w_pass_tc_kernel(const WTcArgs a) {
  using namespace tc;
  if (*a.done) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sW = smem;                                         // W tile [hi 16 KB][lo 16 KB], 128 rows i
  float* sAcc = reinterpret_cast<float*>(smem + WTC_OFF_ACC);       // [32 k][128 lanes]
  float* sQ = reinterpret_cast<float*>(smem + WTC_OFF_Q);           // [4 column slices][128 lanes]
  __shared__ uint64_t bar_w, bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_theta[2], bar_s[2], bar_g;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ib = (int64_t)blockIdx.x * 128;
  const int64_t c0 = (int64_t)blockIdx.y * a.cols_per_split;
  const int64_t c1 = min(a.n, c0 + a.cols_per_split);
  const int nb = c1 > c0 ? (int)((c1 - c0 + 63) / 64) : 0;

  for (int e = tid; e < 32 * 128; e += TC_THREADS) sAcc[e] = 0.0f;
  if (warp == TC_SIMT_WARPS) tmem_alloc(&tmem_base_s, 512);
  if (tid == TC_CTRL_TID) {
    mbar_init(&bar_w, 1);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_theta[0], 1); mbar_init(&bar_theta[1], 1);
    mbar_init(&bar_s[0], TC_SIMT_WARPS); mbar_init(&bar_s[1], TC_SIMT_WARPS);
    mbar_init(&bar_g, 1);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tTheta = tb, tS = tb + 128, tG = tb + 384;
  const int nflush = nb > 0 ? (nb - 1) / kFlush : 0;

  if (tid == TC_CTRL_TID) {
    // ------------------------------------------------------------- control thread
    const int64_t wblk = ib / 64;
    mbar_expect_tx(&bar_w, 32768);
    bulk_g2s(sW, a.f.Wa + (size_t)(wblk * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_w);
    bulk_g2s(sW + 8192, a.f.Wa + (size_t)((wblk + 1) * 2 + 0) * TC_BLK_FLOATS, 8192, &bar_w);
    bulk_g2s(sW + 16384, a.f.Wa + (size_t)(wblk * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_w);
    bulk_g2s(sW + 24576, a.f.Wa + (size_t)((wblk + 1) * 2 + 1) * TC_BLK_FLOATS, 8192, &bar_w);
    auto produce = [&](int b) {
      const int s = b & (TC_STAGES - 1);
      if (b >= TC_STAGES) mbar_wait(&bar_empty[s], ((b / TC_STAGES) - 1) & 1);
      unsigned char* st = smem + WTC_OFF_STAGE + s * WTC_STAGE_BYTES;
      const int64_t hblk = c0 / 64 + b;
      mbar_expect_tx(&bar_full[s], 32768);
      bulk_g2s(st, a.f.Ha + (size_t)hblk * 2 * TC_BLK_FLOATS, 16384, &bar_full[s]);
      bulk_g2s(st + 16384, a.f.Hb + (size_t)hblk * 2 * TC_BLK_FLOATS, 16384, &bar_full[s]);
    };
    const uint64_t dWh = desc_kmajor_sw128(smem_u32(sW)), dWl = desc_kmajor_sw128(smem_u32(sW + 16384));
    constexpr uint32_t id1 = idesc_tf32(128, 64), id2 = idesc_tf32(128, 32);
    auto mma1 = [&](int b) {
      const int s = b & (TC_STAGES - 1);
      mbar_wait(&bar_full[s], (b / TC_STAGES) & 1);
      fence_after_sync();
      unsigned char* st = smem + WTC_OFF_STAGE + s * WTC_STAGE_BYTES;
      const uint64_t dHh = desc_kmajor_sw128(smem_u32(st)), dHl = desc_kmajor_sw128(smem_u32(st + 8192));
      const uint32_t tT = tTheta + 64 * (b & 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tT, dWh + 2 * ks, dHh + 2 * ks, id1, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tT, dWh + 2 * ks, dHl + 2 * ks, id1, 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tT, dWl + 2 * ks, dHh + 2 * ks, id1, 1);
      commit(&bar_theta[b & 1]);
    };
    for (int b = 0; b < nb && b < 3; ++b) produce(b);
    mbar_wait(&bar_w, 0);
    if (nb > 0) mma1(0);
    if (nb > 1) mma1(1);
    for (int b = 0; b < nb; ++b) {
      const int s = b & (TC_STAGES - 1);
      unsigned char* st = smem + WTC_OFF_STAGE + s * WTC_STAGE_BYTES;
      mbar_wait(&bar_s[b & 1], (b >> 1) & 1);
      fence_after_sync();
      const uint32_t tSh = tS + 128 * (b & 1), tSl = tSh + 64;
      const bool chain_start = (b % kFlush) == 0;
#pragma unroll
      for (int t = 0; t < 3; ++t) {                                   // S.H, S.H_lo, S_lo.H
        const uint32_t ta = (t == 2) ? tSl : tSh;
        unsigned char* hb = st + 16384 + (t == 1 ? 8192 : 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dB = desc_kmajor_sw128(smem_u32(hb + (ks >> 2) * 4096)) + 2 * (ks & 3);
          mma_ts(tG, ta + 8 * ks, dB, id2, ((t | ks) > 0 || !chain_start) ? 1u : 0u);
        }
      }
      commit(&bar_empty[s]);
      if (b + 1 == nb || ((b + 1) % kFlush) == 0) commit(&bar_g);
      if (b + 2 < nb) mma1(b + 2);
      if (b + 3 < nb) produce(b + 3);
    }
  } else if (warp < TC_SIMT_WARPS) {
    // ------------------------------------------------------------- SIMT warps: lane = row i, 16 columns j per block
    const int q = warp & 3, sub = warp >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int tl = q * 32 + lane;
    const int64_t row = ib + tl;
    const int64_t rowc = min(row, a.m - 1);
    const uint32_t* __restrict__ Prow = a.P + (size_t)rowc * a.wpr + (c0 >> 5) + (sub >> 1);
    const uint32_t* __restrict__ Mrow = a.M ? a.M + (size_t)rowc * a.wpr + (c0 >> 5) + (sub >> 1) : nullptr;
    const int shift = (sub & 1) * 16;
    const float eps = a.eps;
    float qacc = 0.f;
    auto flush = [&](int idx) {
      mbar_wait(&bar_g, idx & 1);
      fence_after_sync();
      uint32_t v[8];
      tmem_ld8(tG + lane_off + 8 * sub, v);
      wait_ld();
#pragma unroll
      for (int e = 0; e < 8; ++e) sAcc[(8 * sub + e) * 128 + tl] += __uint_as_float(v[e]);
    };
    for (int b = 0; b < nb; ++b) {
      const int64_t colw = c0 + 64 * (int64_t)b + 32 * (sub >> 1);    // first column of this thread's bit word
      const int64_t rem = c1 - colw;
      const uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
      const uint32_t pbits = Prow[2 * b] >> shift;
      const uint32_t mbits = ((Mrow ? Mrow[2 * b] : 0xffffffffu) & valid) >> shift;
      if (b > 0 && (b % kFlush) == 0) flush(b / kFlush - 1);
      mbar_wait(&bar_theta[b & 1], (b >> 1) & 1);
      fence_after_sync();
      uint32_t v[16], sh[16], sl[16];
      tmem_ld16(tTheta + 64 * (b & 1) + lane_off + 16 * sub, v);
      wait_ld();
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
      const uint32_t tSh = tS + 128 * (b & 1) + lane_off + 16 * sub;
      tmem_st16(tSh, sh);
      tmem_st16(tSh + 64, sl);
      wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s[b & 1]);
    }
    if (nb > 0) flush(nflush);
    sQ[sub * 128 + tl] = qacc;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == TC_SIMT_WARPS) tmem_dealloc(tb, 512);
  if (warp < TC_SIMT_WARPS) {
    const int q = warp & 3, sub = warp >> 2, tl = q * 32 + lane;
    const int64_t row = ib + tl;
    if (row < a.m) {
      float* __restrict__ Gg = a.G + ((size_t)blockIdx.y * a.m + row) * 32 + 8 * sub;
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = sAcc[(8 * sub + e) * 128 + tl];
      *reinterpret_cast<float4*>(Gg) = make_float4(g[0], g[1], g[2], g[3]);
      *reinterpret_cast<float4*>(Gg + 4) = make_float4(g[4], g[5], g[6], g[7]);
      if (sub == 0)                                                   // fixed order: bit-reproducible
        a.Q[(size_t)blockIdx.y * a.m + row] = ((sQ[tl] + sQ[128 + tl]) + sQ[256 + tl]) + sQ[384 + tl];
    }
  }
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

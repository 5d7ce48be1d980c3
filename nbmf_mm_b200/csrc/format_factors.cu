// Operand formatting for the tensor-core pass kernels (tc_passes.cuh).
//
// Factors: every streamed factor block is split into tf32 hi + lo parts and written in the K-major,
// 128-byte-swizzled layout the tcgen05 shared-memory descriptors expect, all operand forms of one
// block contiguous, so a pipeline stage of the hot kernels is ONE 1-D bulk copy.
//   Wf [mpad/32][4][KT/32 x 4 KB]  per 32-row block of W:    rows i x k  tf32 hi | bf16 correction  (B of the H pass MMA1)
//                                                           rows k x i  tf32 hi | bf16 correction  (B of the H pass MMA2)
//   Hf [ldh/64][4][KT/32 x 8 KB]   per 64-column block of H: rows j x k tf32 hi | bf16 correction  (B of the W pass MMA1)
//                                                           2 K-blocks of rows k x 32 j, tf32 hi | bf16 correction (MMA2)
//   (KT = 32 for K <= 32, 64 for K <= 64)
// Every product a.b = hi.hi + (hi.lo + lo.hi): the first term is a TF32 MMA chain, the two correction terms come
// from ONE bf16 MMA chain with twice the K extent.  MMA1: the A tile holds [hi | lo] as bf16, the streamed
// correction plane [lo | hi] per row.  MMA2: the SIMT stage packs (hi, lo) of a ratio into one 32-bit column
// (cvt.rn.bf16x2), the streamed plane holds the matching (lo, hi) pairs.  bf16 keeps the fp32 exponent, and 8
// bits are enough for terms that are 2^-11 of the sum.
// Bit planes: the SIMT threads of the tensor kernels own one TMEM lane each (a column j in the H pass, a
// row i in the W pass) and walk along the other axis, so the planes are re-tiled once per fit so that a
// warp's 32 lanes read 32 consecutive words:
//   Pc [ldh/128][mpad/32][128]        word = rows 32 rb .. 32 rb + 31 of column 128 jt + jj of P
//   PM [mpad/128][wpr][128] (uint2)   {P word, observed word} of row 128 it + ii, columns 32 cw .. 32 cw + 31
#include <algorithm>

#include "internal.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace nbmf {

// KT = 32 (K <= 32) or 64 (K <= 64): K extent of the formatted operands.  A K-major operand wider than 128 bytes per row is
// stored as KT/32 slabs (one SWIZZLE_128B tile per 32 k / per 64 bf16), slab after slab.
//
// Both kernels stage the source block in shared memory and then walk the DESTINATION linearly (thread t writes 32-bit word
// t, t + 256, ..: full coalesced lines), inverting the swizzle to find the source element -- the first version walked the
// source and scattered 2- and 4-byte stores into the transposed regions (67 us per iteration for a batch of 64 small fits,
// more than the H pass itself; 0.8 ms per iteration at config 4).
__device__ __forceinline__ void unswizzle(uint32_t byte_off, int& row, int& chunk, int& within) {
  // inverse of sw128_offset*: byte offset inside a tile of 128-byte rows -> (row, logical 16-byte chunk, byte in chunk)
  row = (int)((byte_off >> 10) << 3) + (int)((byte_off & 1023u) >> 7);
  chunk = (int)(((byte_off & 127u) >> 4) ^ (uint32_t)(row & 7));
  within = (int)(byte_off & 15u);
}

template <int KT>
__global__ void __launch_bounds__(256) format_w_kernel(const float* __restrict__ W, int64_t m, int64_t mpad, float* __restrict__ Wf,
                                                       const FitState* __restrict__ state, int64_t bstride) {
  if (bstride) {                                  // fit blockIdx.z of a batch: its own workspace
    const size_t sh = (size_t)blockIdx.z * (size_t)bstride;
    W = batch_shift(W, sh); Wf = batch_shift(Wf, sh);
    if (state) state = batch_shift(state, sh);
  }
  if (state && state->done) return;
  constexpr int REGF = (KT / 32) * 1024;                           // 32-bit words per operand region of a 32-row block
  __shared__ float tile[32][KT + 1];
  const int64_t rb = blockIdx.x;
  for (int e = threadIdx.x; e < 32 * KT; e += 256) {
    const int r = e / KT, k = e % KT;
    const int64_t row = rb * 32 + r;
    tile[r][k] = row < m ? W[row * KT + k] : 0.0f;
  }
  __syncthreads();
  uint32_t* blk = reinterpret_cast<uint32_t*>(Wf + (size_t)rb * (4 * REGF));
  for (int d = threadIdx.x; d < REGF; d += 256) {
    int r, c, w;
    unswizzle((uint32_t)(d & 1023) * 4u, r, c, w);
    const int slab = d >> 10;
    // region 0: rows i x k, tf32 hi (slab = 32 k)
    blk[d] = __float_as_uint(tc::tf32_trunc(tile[r][32 * slab + 4 * c + (w >> 2)]));
    // region 1: rows i: [lo (KT) | hi (KT)] bf16, two per word (slab = 64 bf16)
    {
      uint32_t word = 0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = 64 * slab + 8 * c + (w >> 1) + q;
        const float x = tile[r][idx < KT ? idx : idx - KT];
        const float hi = tc::tf32_trunc(x);
        word |= tc::bf16_bits(idx < KT ? x - hi : hi) << (16 * q);
      }
      blk[REGF + d] = word;
    }
    // regions 2, 3: rows k (KT of them, 128 bytes each) x i: tf32 hi, and (lo, hi) bf16 pairs per i
    {
      int k, ci, wi;
      unswizzle((uint32_t)d * 4u, k, ci, wi);
      blk[2 * REGF + d] = __float_as_uint(tc::tf32_trunc(tile[4 * ci + (wi >> 2)][k]));
      const float x = tile[(8 * ci + (wi >> 1)) >> 1][k];
      const float hi = tc::tf32_trunc(x);
      blk[3 * REGF + d] = tc::bf16_bits(x - hi) | (tc::bf16_bits(hi) << 16);
    }
  }
}

template <int KT>
__global__ void __launch_bounds__(256) format_h_kernel(const float* __restrict__ H, int64_t ldh, float* __restrict__ Hf,
                                                       const FitState* __restrict__ state, int64_t bstride, float theta_bias) {
  if (bstride) {
    const size_t sh = (size_t)blockIdx.z * (size_t)bstride;
    H = batch_shift(H, sh); Hf = batch_shift(Hf, sh);
    if (state) state = batch_shift(state, sh);
  }
  if (state && state->done) return;
  constexpr int RWF = (KT / 32) * 2048;                            // words per operand region of a 64-column block
  constexpr int RBF = KT * 32;                                     // words per K-block (32 columns) of the H operand
  __shared__ float tile[KT][64 + 1];
  const int64_t jb = blockIdx.x;
  for (int e = threadIdx.x; e < KT * 64; e += 256) {
    const int k = e >> 6, c = e & 63;
    tile[k][c] = H[(size_t)k * ldh + jb * 64 + c];
  }
  __syncthreads();
  uint32_t* blk = reinterpret_cast<uint32_t*>(Hf + (size_t)jb * (4 * RWF));
  for (int d = threadIdx.x; d < RWF; d += 256) {
    {  // regions 0, 1: rows j (64) x k: tf32 hi (slab = 32 k) and [lo (KT) | hi (KT)] bf16 (slab = 64 bf16)
      int r, c, w;
      unswizzle((uint32_t)(d & 2047) * 4u, r, c, w);
      const int slab = d >> 11;
      // theta_bias: W's rows sum to one, so Theta + eps = sum_k W_ik (H_kj + eps) -- the MMA1 operand of the K <= 32 W
      // pass is formatted from H + eps and the ones' x is the MMA output itself (one FADD per entry less, w_entry)
      blk[d] = __float_as_uint(tc::tf32_trunc(tile[32 * slab + 4 * c + (w >> 2)][r] + theta_bias));
      uint32_t word = 0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = 64 * slab + 8 * c + (w >> 1) + q;
        const float x = tile[idx < KT ? idx : idx - KT][r] + theta_bias;
        const float hi = tc::tf32_trunc(x);
        word |= tc::bf16_bits(idx < KT ? x - hi : hi) << (16 * q);
      }
      blk[RWF + d] = word;
    }
    {  // regions 2, 3: two K-blocks of rows k x 32 j: tf32 hi, and (lo, hi) bf16 pairs per j
      const int hb = d / RBF;
      int k, cj, wj;
      unswizzle((uint32_t)(d % RBF) * 4u, k, cj, wj);
      blk[2 * RWF + d] = __float_as_uint(tc::tf32_trunc(tile[k][32 * hb + 4 * cj + (wj >> 2)]));
      const float x = tile[k][32 * hb + ((8 * cj + (wj >> 1)) >> 1)];
      const float hi = tc::tf32_trunc(x);
      blk[3 * RWF + d] = tc::bf16_bits(x - hi) | (tc::bf16_bits(hi) << 16);
    }
  }
}

void launch_format_w(const void* W, int64_t m, int64_t mpad, int kt, void* Wf, const FitState* state, cudaStream_t st,
                     int batch_n, int64_t bstride) {
  const dim3 grid((unsigned)(mpad / 32), 1, (unsigned)batch_n);       // one block per 32-row block of W
  if (kt == 64) format_w_kernel<64><<<grid, 256, 0, st>>>((const float*)W, m, mpad, (float*)Wf, state, bstride);
  else format_w_kernel<32><<<grid, 256, 0, st>>>((const float*)W, m, mpad, (float*)Wf, state, bstride);
}
void launch_format_h(const void* H, int64_t ldh, int kt, float theta_bias, void* Hf, const FitState* state, cudaStream_t st,
                     int batch_n, int64_t bstride) {
  dim3 grid((unsigned)(ldh / 64), 1, (unsigned)batch_n);              // one block per 64-column block of H
  if (kt == 64) format_h_kernel<64><<<grid, 256, 0, st>>>((const float*)H, ldh, (float*)Hf, state, bstride, theta_bias);
  else format_h_kernel<32><<<grid, 256, 0, st>>>((const float*)H, ldh, (float*)Hf, state, bstride, theta_bias);
}

// Ones per column of a re-tiled plane (Pc layout), accumulated over the row blocks [rb0, rb1): the H pass uses the
// density of ones of a column to pick the plane it accumulates directly.  Integer atomics: deterministic.
__global__ void colcount_kernel(const uint32_t* __restrict__ Pc, int64_t nrb, int64_t rb0, int64_t rb1, int64_t ldh,
                                uint32_t* __restrict__ colcnt) {
  const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;       // one column block of 128 per CTA.x
  if (j >= ldh) return;
  const int64_t per = (rb1 - rb0 + gridDim.y - 1) / gridDim.y;
  const int64_t a0 = rb0 + (int64_t)blockIdx.y * per, a1 = min(rb1, a0 + per);
  const uint32_t* __restrict__ p = Pc + ((size_t)blockIdx.x * nrb) * 128 + threadIdx.x;
  uint32_t c = 0;
  for (int64_t rb = a0; rb < a1; ++rb) c += __popc(p[(size_t)rb * 128]);
  if (c) atomicAdd(&colcnt[j], c);
}
void launch_colcount(const uint32_t* Pc, int64_t nrb, int64_t rb0, int64_t rb1, int64_t ldh, uint32_t* colcnt, cudaStream_t st) {
  if (rb1 <= rb0) return;
  const int64_t ncb = ldh / 128;
  int ysplit = (int)std::min<int64_t>(64, std::max<int64_t>(1, (148 * 8 + ncb - 1) / ncb));
  ysplit = (int)std::min<int64_t>(ysplit, rb1 - rb0);
  dim3 grid((unsigned)ncb, (unsigned)ysplit);
  colcount_kernel<<<grid, 128, 0, st>>>(Pc, nrb, rb0, rb1, ldh, colcnt);
}

// One warp per 32 x 32 bit tile: lane r holds the word of row 32 rb + r, 32 ballots transpose it, lane c
// writes the word of column 32 cw + c.  The 8 warps of a block take 8 adjacent column words, so the
// strided row reads share their 32-byte sectors and every write is a full 128-byte line.
__global__ void tile_pc_kernel(const uint32_t* __restrict__ P, int64_t m, int64_t wpr, int64_t nrb, int64_t rb0,
                               uint32_t* __restrict__ Pc) {
  const int lane = threadIdx.x & 31;
  const int64_t cw = (int64_t)blockIdx.y * 8 + (threadIdx.x >> 5);
  const int64_t rb = rb0 + blockIdx.x;
  if (cw >= wpr) return;
  const int64_t row = rb * 32 + lane;
  const uint32_t w = row < m ? P[row * wpr + cw] : 0u;
  uint32_t mine = 0;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    const uint32_t t = __ballot_sync(0xffffffffu, (w >> b) & 1u);
    if (lane == b) mine = t;
  }
  Pc[((size_t)(cw >> 2) * nrb + rb) * 128 + (size_t)(cw & 3) * 32 + lane] = mine;
}

// Block = 128 rows x 8 column words: read with 8 consecutive words per row (one sector), written with 128
// consecutive rows per column word.
__global__ void tile_pm_kernel(const uint32_t* __restrict__ P, const uint32_t* __restrict__ M, int64_t m, int64_t n,
                               int64_t wpr, int64_t it0, uint2* __restrict__ PM) {
  __shared__ uint2 tile[8][129];
  const int64_t it = it0 + blockIdx.x, cw0 = (int64_t)blockIdx.y * 8;
  for (int t = threadIdx.x; t < 1024; t += blockDim.x) {
    const int ii = t >> 3, c = t & 7;
    const int64_t row = it * 128 + ii, cw = cw0 + c;
    uint2 v = make_uint2(0u, 0u);
    if (row < m && cw < wpr) {
      const int64_t rem = n - cw * 32;
      const uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
      v.x = P[row * wpr + cw] & valid;
      v.y = (M ? M[row * wpr + cw] : 0xffffffffu) & valid;
    }
    tile[c][ii] = v;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 1024; t += blockDim.x) {
    const int c = t >> 7, ii = t & 127;
    if (cw0 + c < wpr) PM[((size_t)it * wpr + cw0 + c) * 128 + ii] = tile[c][ii];
  }
}

// Rows [row0, row1) of the planes (row0 a multiple of 128; the chunk that ends at m also writes the zero padding up to
// mpad): the whole matrix at once, or chunk by chunk while later rows are still crossing PCIe.
void launch_tile_planes(const uint32_t* P, const uint32_t* M, int64_t m, int64_t n, int64_t wpr, int64_t mpad,
                        int64_t row0, int64_t row1, uint32_t* Pc, uint32_t* Mc, void* PM, cudaStream_t st) {
  const int64_t nrb = mpad / 32;
  const int64_t end = row1 >= m ? mpad : row1;
  if (end <= row0) return;
  dim3 g1((unsigned)((end - row0 + 31) / 32), (unsigned)((wpr + 7) / 8));
  tile_pc_kernel<<<g1, 256, 0, st>>>(P, m, wpr, nrb, row0 / 32, Pc);
  if (Mc && M) tile_pc_kernel<<<g1, 256, 0, st>>>(M, m, wpr, nrb, row0 / 32, Mc);   // strict mask semantics: the H pass needs M too
  dim3 g2((unsigned)((end - row0 + 127) / 128), (unsigned)((wpr + 7) / 8));
  tile_pm_kernel<<<g2, 256, 0, st>>>(P, M, m, n, wpr, row0 / 128, (uint2*)PM);
}

}  // namespace nbmf

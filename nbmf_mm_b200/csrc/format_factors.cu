// Operand formatting for the tensor-core pass kernels (tc_passes.cuh).
//
// Factors: every streamed factor block is split into tf32 hi + lo parts and written in the K-major,
// 128-byte-swizzled layout the tcgen05 shared-memory descriptors expect, all operand forms of one
// block contiguous, so a pipeline stage of the hot kernels is ONE 1-D bulk copy.
//   Wf [mpad/32][4][4 KB]   per 32-row block of W:   rows i x k  tf32 hi | bf16 correction  (B of the H pass MMA1)
//                                                    rows k x i  tf32 hi | bf16 correction  (B of the H pass MMA2)
//   Hf [ldh/64][4][8 KB]    per 64-column block of H: rows j x k tf32 hi | bf16 correction  (B of the W pass MMA1)
//                                                    2 K-blocks of rows k x 32 j, tf32 hi | bf16 correction (MMA2)
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
#include "internal.h"
#include "tc_common.cuh"

namespace nbmf {

__global__ void format_w_kernel(const float* __restrict__ W, int64_t m, int64_t mpad, float* __restrict__ Wf,
                                const FitState* __restrict__ state) {
  if (state && state->done) return;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= mpad * 32) return;
  const int64_t i = e >> 5;
  const int k = (int)(e & 31);
  const float x = i < m ? W[e] : 0.0f;
  const float hi = tc::tf32_trunc(x), lo = x - hi;
  const int r = (int)(i & 31);
  float* blk = Wf + (size_t)(i >> 5) * 4096;
  const uint32_t oa = tc::sw128_offset(r, k) / 4, ob = tc::sw128_offset(k, r) / 4;
  blk[oa] = hi;
  unsigned short* corr = reinterpret_cast<unsigned short*>(blk + 1024);
  corr[tc::sw128_offset_b16(r, k) / 2] = (unsigned short)tc::bf16_bits(lo);
  corr[tc::sw128_offset_b16(r, 32 + k) / 2] = (unsigned short)tc::bf16_bits(hi);
  blk[2048 + ob] = hi;
  unsigned short* corr2 = reinterpret_cast<unsigned short*>(blk + 3072);     // row k: (lo, hi) pairs per i
  corr2[tc::sw128_offset_b16(k, 2 * r) / 2] = (unsigned short)tc::bf16_bits(lo);
  corr2[tc::sw128_offset_b16(k, 2 * r + 1) / 2] = (unsigned short)tc::bf16_bits(hi);
}

__global__ void format_h_kernel(const float* __restrict__ H, int64_t ldh, float* __restrict__ Hf,
                                const FitState* __restrict__ state) {
  if (state && state->done) return;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (j >= ldh) return;
  const float x = H[(size_t)k * ldh + j];
  const float hi = tc::tf32_trunc(x), lo = x - hi;
  const int r = (int)(j & 63);
  float* blk = Hf + (size_t)(j >> 6) * 8192;
  const uint32_t oa = tc::sw128_offset(r, k) / 4;
  const uint32_t ob = (uint32_t)(r >> 5) * 1024 + tc::sw128_offset(k, r & 31) / 4;
  blk[oa] = hi;
  unsigned short* corr = reinterpret_cast<unsigned short*>(blk + 2048);
  corr[tc::sw128_offset_b16(r, k) / 2] = (unsigned short)tc::bf16_bits(lo);
  corr[tc::sw128_offset_b16(r, 32 + k) / 2] = (unsigned short)tc::bf16_bits(hi);
  blk[4096 + ob] = hi;
  unsigned short* corr2 = reinterpret_cast<unsigned short*>(blk + 6144);     // 2 K-blocks; row k: (lo, hi) pairs per j
  const int rr = r & 31;
  corr2[((r >> 5) * 4096 + tc::sw128_offset_b16(k, 2 * rr)) / 2] = (unsigned short)tc::bf16_bits(lo);
  corr2[((r >> 5) * 4096 + tc::sw128_offset_b16(k, 2 * rr + 1)) / 2] = (unsigned short)tc::bf16_bits(hi);
}

void launch_format_w(const void* W, int64_t m, int64_t mpad, void* Wf, const FitState* state, cudaStream_t st) {
  const unsigned grid = (unsigned)((mpad * 32 + 255) / 256);
  format_w_kernel<<<grid, 256, 0, st>>>((const float*)W, m, mpad, (float*)Wf, state);
}
void launch_format_h(const void* H, int64_t ldh, void* Hf, const FitState* state, cudaStream_t st) {
  dim3 grid((unsigned)((ldh + 255) / 256), 32);
  format_h_kernel<<<grid, 256, 0, st>>>((const float*)H, ldh, (float*)Hf, state);
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

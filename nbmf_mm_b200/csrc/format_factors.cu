// Operand formatting for the tensor-core pass kernels (tc_passes.cuh): every factor is split into
// tf32 hi + lo parts and written in the K-major, 128-byte-swizzled block layout the tcgen05 shared
// memory descriptors expect, so the hot kernels fill their stages with plain 1-D bulk copies.
//   Wa [mpad/64][hi|lo][64 rows i][32 k]          W blocks   (A of the W pass, B of the H pass MMA1)
//   Wb [mpad/64][hi|lo][2][32 rows k][32 i]       W^T blocks (B of the H pass MMA2)
//   Ha [ldh/64][hi|lo][64 rows j][32 k]           Ht blocks  (A of the H pass, B of the W pass MMA1)
//   Hb [ldh/64][hi|lo][2][32 rows k][32 j]        H blocks   (B of the W pass MMA2)
#include "internal.h"
#include "tc_common.cuh"

namespace nbmf {

__global__ void format_w_kernel(const float* __restrict__ W, int64_t m, int64_t mpad, float* __restrict__ Wa,
                                float* __restrict__ Wb, const FitState* __restrict__ state) {
  if (state && state->done) return;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= mpad * 32) return;
  const int64_t i = e >> 5;
  const int k = (int)(e & 31);
  const float x = i < m ? W[e] : 0.0f;
  const float hi = tc::tf32_trunc(x), lo = x - hi;
  const int64_t blk = i >> 6;
  const int r = (int)(i & 63);
  const size_t oa = (size_t)blk * 2 * 2048 + tc::sw128_offset(r, k) / 4;
  Wa[oa] = hi;
  Wa[oa + 2048] = lo;
  const size_t ob = (size_t)blk * 2 * 2048 + (size_t)(r >> 5) * 1024 + tc::sw128_offset(k, r & 31) / 4;
  Wb[ob] = hi;
  Wb[ob + 2048] = lo;
}

__global__ void format_h_kernel(const float* __restrict__ H, int64_t ldh, float* __restrict__ Ha,
                                float* __restrict__ Hb, const FitState* __restrict__ state) {
  if (state && state->done) return;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (j >= ldh) return;
  const float x = H[(size_t)k * ldh + j];
  const float hi = tc::tf32_trunc(x), lo = x - hi;
  const int64_t blk = j >> 6;
  const int r = (int)(j & 63);
  const size_t oa = (size_t)blk * 2 * 2048 + tc::sw128_offset(r, k) / 4;
  Ha[oa] = hi;
  Ha[oa + 2048] = lo;
  const size_t ob = (size_t)blk * 2 * 2048 + (size_t)(r >> 5) * 1024 + tc::sw128_offset(k, r & 31) / 4;
  Hb[ob] = hi;
  Hb[ob + 2048] = lo;
}

void launch_format_w(const void* W, int64_t m, int64_t mpad, void* Wa, void* Wb, const FitState* state, cudaStream_t st) {
  const unsigned grid = (unsigned)((mpad * 32 + 255) / 256);
  format_w_kernel<<<grid, 256, 0, st>>>((const float*)W, m, mpad, (float*)Wa, (float*)Wb, state);
}
void launch_format_h(const void* H, int64_t ldh, void* Ha, void* Hb, const FitState* state, cudaStream_t st) {
  dim3 grid((unsigned)((ldh + 255) / 256), 32);
  format_h_kernel<<<grid, 256, 0, st>>>((const float*)H, ldh, (float*)Ha, (float*)Hb, state);
}

}  // namespace nbmf

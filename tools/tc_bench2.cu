// tcgen05.mma throughput under TMEM / shared-memory contention (development experiment).
// One warp issues TS-form kind::tf32 MMAs (M=128, N=32, K=8) back to back; `noise` other warps run
// tcgen05.ld / tcgen05.st loops (mode 1, 2) or MUFU loops (mode 3) at the same time.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../nbmf_mm_b200/csrc/tc_common.cuh"
using namespace nbmf::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(672) k(int mode, int noise_warps, int count, long long* out, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int e = tid; e < 32768 / 4; e += blockDim.x) reinterpret_cast<float*>(smem)[e] = 0.001f * (e & 31);
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tb = tmem_base_s;
  if (warp == 20) {
    const bool leader = elect_one();
    const uint32_t id = idesc_tf32(128, 32);
    const uint64_t dB = desc_kmajor_sw128(smem_u32(smem));
    const long long t0 = clock64();
    for (int i = 0; i < count; i += 8) if (leader) {
      mma_ts(tb, tb + 256, dB, id, 1); mma_ts(tb + 32, tb + 264, dB + 2, id, 1); mma_ts(tb, tb + 272, dB + 4, id, 1); mma_ts(tb + 32, tb + 280, dB + 6, id, 1);
      mma_ts(tb, tb + 256, dB, id, 1); mma_ts(tb + 32, tb + 264, dB + 2, id, 1); mma_ts(tb, tb + 272, dB + 4, id, 1); mma_ts(tb + 32, tb + 280, dB + 6, id, 1);
    }
    if (leader) commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (leader) { out[0] = t1 - t0; }
    stop = 1;
  } else if (warp < noise_warps) {
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col = tb + 320 + 32 * ((warp >> 2) & 3) + lane_off;   // columns the MMAs do not touch
    uint32_t v[32];
    for (int e = 0; e < 32; ++e) v[e] = tid + e;
    float acc = 0.f;
    long long n = 0;
    while (!stop) {
      if (mode == 1) { uint32_t w[8]; tmem_ld8(col, w); wait_ld(); acc += __uint_as_float(w[0]); }
      else if (mode == 2) { tmem_st32(col, v); wait_st(); }
      else if (mode == 3) { for (int e = 0; e < 8; ++e) acc += __frcp_rn(acc + (float)e); }
      else if (mode == 4) { uint32_t w[8]; tmem_ld8(col, w); wait_ld(); acc += __uint_as_float(w[0]); tmem_st32(col, v); wait_st(); }
      ++n;
    }
    if (lane == 0 && warp == 0) out[1] = n;
    if (acc == 12345.f) sink[0] = acc;
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d_out; float* sink;
  CK(cudaMalloc(&d_out, 16)); CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
  const int count = 8192;
  const char* names[] = {"idle", "ld8 loop", "st32 loop", "mufu loop", "ld8+st32 loop"};
  for (int mode = 0; mode <= 4; ++mode)
    for (int nw : {0, 4, 8, 16}) {
      if ((mode == 0) != (nw == 0)) continue;
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 672, 32768 + 1024>>>(mode, nw, count, d_out, sink); CK(cudaDeviceSynchronize()); }
      long long h[2];
      CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
      printf("noise %-14s warps %2d : %6.1f cycles/MMA   (noise iterations of warp 0: %lld, %.1f cycles each)\n", names[mode], nw,
             (double)h[0] / count, h[1], h[1] ? (double)h[0] / h[1] : 0.0);
    }
  return 0;
}

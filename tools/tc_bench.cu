// tcgen05.mma issue/throughput microbenchmark (round-1 experiment): cycles per MMA as a function of N,
// operand form (SS / TS) and the number of independent TMEM accumulators the issue loop rotates over.
// Answers: how expensive is a chain of small dependent kind::tf32 MMAs (M=128, N=32, K=8)?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

// mode 0: SS, mode 1: TS.  nacc independent accumulators of N columns each (nacc * N <= 256).
__global__ void __launch_bounds__(128) bench_kernel(int mode, int N, int nacc, int count, long long* out, int issuers) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  for (int e = tid; e < 49152 / 4; e += 128) reinterpret_cast<float*>(smem)[e] = 0.001f * (e & 31);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(issuers));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  // The whole issuing warp runs this region (uniform control flow, operands in uniform registers);
  // only the elected lane executes the MMAs.  Issuing from a divergent `if (tid == x)` makes the
  // compiler wrap every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~54 cycles per MMA).
  if (warp < issuers) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(leader));
    const uint32_t id = idesc_tf32(128, N);
    const uint32_t tb = tmem_base_s + (uint32_t)(warp * 64);   // each issuer has its own 64 accumulator columns
    const uint64_t dA = desc_kmajor_sw128(smem_u32(smem)), dB = desc_kmajor_sw128(smem_u32(smem + 16384));
    const uint32_t tA = tmem_base_s + 256;                  // A operand region for TS (contents irrelevant)
    const long long t0 = clock64();
    // unrolled by 8 with operands hoisted: the loop body is nothing but 8 MMA issues
    const uint32_t d0 = tb, d1 = tb + (1 % nacc) * N, d2 = tb + (2 % nacc) * N, d3 = tb + (3 % nacc) * N;
    const uint32_t d4 = tb + (4 % nacc) * N, d5 = tb + (5 % nacc) * N, d6 = tb + (6 % nacc) * N, d7 = tb + (7 % nacc) * N;
    if (mode == 0) {
      for (int i = 0; i < count; i += 8) if (leader) {
        mma_ss(d0, dA, dB, id, 1); mma_ss(d1, dA + 2, dB + 2, id, 1); mma_ss(d2, dA + 4, dB + 4, id, 1); mma_ss(d3, dA + 6, dB + 6, id, 1);
        mma_ss(d4, dA, dB, id, 1); mma_ss(d5, dA + 2, dB + 2, id, 1); mma_ss(d6, dA + 4, dB + 4, id, 1); mma_ss(d7, dA + 6, dB + 6, id, 1);
      }
    } else {
      for (int i = 0; i < count; i += 8) if (leader) {
        mma_ts(d0, tA, dB, id, 1); mma_ts(d1, tA + 8, dB + 2, id, 1); mma_ts(d2, tA + 16, dB + 4, id, 1); mma_ts(d3, tA + 24, dB + 6, id, 1);
        mma_ts(d4, tA, dB, id, 1); mma_ts(d5, tA + 8, dB + 2, id, 1); mma_ts(d6, tA + 16, dB + 4, id, 1); mma_ts(d7, tA + 24, dB + 6, id, 1);
      }
    }
    const long long t1 = clock64();
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    if (warp == 0)
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0 && warp == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 16));
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
  const int count = 4096;
  printf("%-4s %-5s %-7s %16s %24s\n", "form", "N", "issuers", "issue cyc/mma", "aggregate cyc/mma (all)");
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 32, 64})
      for (int issuers : {1, 2, 3, 4}) {
        for (int rep = 0; rep < 2; ++rep) {
          bench_kernel<<<1, 128, 49152 + 1024>>>(mode, N, 1, count, d_out, issuers);
          CK(cudaDeviceSynchronize());
        }
        long long h[2];
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("%-4s %-5d %-7d %16.1f %24.1f\n", mode ? "TS" : "SS", N, issuers, (double)h[0] / count, (double)h[1] / (count * issuers));
      }
  // one issuer rotating over nacc independent accumulators: separates the issue cost from the
  // latency of a dependent accumulate chain
  printf("%-4s %-5s %-7s %16s %24s\n", "form", "N", "nacc", "issue cyc/mma", "total cyc/mma");
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 32, 64, 128, 256})
      for (int nacc : {1, 2, 4, 8}) {
        if (nacc * N > 256) continue;
        for (int rep = 0; rep < 2; ++rep) {
          bench_kernel<<<1, 128, 49152 + 1024>>>(mode, N, nacc, count, d_out, 1);
          CK(cudaDeviceSynchronize());
        }
        long long h[2];
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("%-4s %-5d %-7d %16.1f %24.1f\n", mode ? "TS" : "SS", N, nacc, (double)h[0] / count, (double)h[1] / count);
      }
  return 0;
}

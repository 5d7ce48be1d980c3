// tcgen05 bring-up probe (round-1 experiment, not part of the product path).
//   test A (SS): D[128 x N1] = A[128 x 32] . B[N1 x 32]^T, tf32, both operands K-major in shared memory
//                with the 128-byte swizzle written by hand (no TMA), 1-term and 3-term (hi/lo) products.
//   test B (TS): D2[128 x 32] = S[128 x 64] . B2[32 x 64]^T with S written to TMEM by tcgen05.st.
// Both are checked against an fp64 host reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);      // start address >> 4
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool wait_parity(uint64_t* bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
#define TMEM_LD16(addr, v, o)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),       \
                 "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),    \
                 "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                  \
               : "r"(addr))
#define TMEM_ST16(addr, v, o)                                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(addr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),   \
                 "r"(v[o + 6]), "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]),         \
                 "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]) : "memory")

// rows x 32 fp32 (128 B per row) -> K-major SWIZZLE_128B tile: 16-byte chunk c of row r lands at chunk c ^ (r & 7)
__device__ void fill_tile(float* tile, const float* src, int rows, int ld, int tid, int nt) {
  for (int e = tid; e < rows * 8; e += nt) {
    const int r = e >> 3, c = e & 7;
    const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * ld + 4 * c);
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(tile) + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

constexpr int N1 = 64;

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    const float* __restrict__ S, const float* __restrict__ B2,
                                                    float* __restrict__ D1, float* __restrict__ D3,
                                                    float* __restrict__ D2, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);                    // 128 x 32   (16 KB)
  float* sAl = reinterpret_cast<float*>(smem + 16384);           // lo part
  float* sB = reinterpret_cast<float*>(smem + 32768);            // 64 x 32    (8 KB)
  float* sBl = reinterpret_cast<float*>(smem + 40960);
  float* sB2 = reinterpret_cast<float*>(smem + 49152);           // 2 sub-tiles of 32 x 32 (4 KB each)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // hi operands are the raw fp32 values (the tensor core reads the tf32 bits), lo = x - tf32(x)
  fill_tile(sA, A, 128, 32, tid, 128);
  fill_tile(sB, B, N1, 32, tid, 128);
  for (int e = tid; e < 128 * 32; e += 128) {
    const int r = e >> 5, k = e & 31;
    const float x = A[e];
    reinterpret_cast<float*>(reinterpret_cast<char*>(sAl) + r * 128 + (((k >> 2) ^ (r & 7)) << 4))[k & 3] = x - tf32_hi(x);
  }
  for (int e = tid; e < N1 * 32; e += 128) {
    const int r = e >> 5, k = e & 31;
    const float x = B[e];
    reinterpret_cast<float*>(reinterpret_cast<char*>(sBl) + r * 128 + (((k >> 2) ^ (r & 7)) << 4))[k & 3] = x - tf32_hi(x);
  }
  // B2: 32 rows (n) x 64 (k) as two K-blocks of 32
  fill_tile(sB2, B2, 32, 64, tid, 128);
  fill_tile(sB2 + 32 * 32, B2 + 32, 32, 64, tid, 128);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t tD1 = tbase, tD3 = tbase + 64, tS = tbase + 128, tD2 = tbase + 192;

  // ---- test A: 1-term into tD1, 3-term into tD3
  if (tid == 0) {
    const uint32_t id = idesc_tf32(128, N1);
    const uint64_t dA = desc_kmajor_sw128(smem_u32(sA)), dAl = desc_kmajor_sw128(smem_u32(sAl));
    const uint64_t dB = desc_kmajor_sw128(smem_u32(sB)), dBl = desc_kmajor_sw128(smem_u32(sBl));
    for (int ks = 0; ks < 4; ++ks) mma_ss(tD1, dA + 2 * ks, dB + 2 * ks, id, ks > 0);
    for (int ks = 0; ks < 4; ++ks) mma_ss(tD3, dA + 2 * ks, dB + 2 * ks, id, ks > 0);
    for (int ks = 0; ks < 4; ++ks) mma_ss(tD3, dA + 2 * ks, dBl + 2 * ks, id, 1);
    for (int ks = 0; ks < 4; ++ks) mma_ss(tD3, dAl + 2 * ks, dB + 2 * ks, id, 1);
    commit(&bar);
  }
  // ---- meanwhile: every thread writes its row of S (64 values) into TMEM columns tS .. tS+63
  {
    uint32_t v[64];
    for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(S[(size_t)tid * 64 + j]);
    const uint32_t addr = tS + ((uint32_t)(warp * 32) << 16);
    TMEM_ST16(addr, v, 0);
    TMEM_ST16(addr + 16, v, 16);
    TMEM_ST16(addr + 32, v, 32);
    TMEM_ST16(addr + 48, v, 48);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (!wait_parity(&bar, 0)) { if (tid == 0) status[0] = 1; return; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t v[64];
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    TMEM_LD16(tD1 + lane_off, v, 0); TMEM_LD16(tD1 + lane_off + 16, v, 16);
    TMEM_LD16(tD1 + lane_off + 32, v, 32); TMEM_LD16(tD1 + lane_off + 48, v, 48);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 64; ++j) D1[(size_t)tid * N1 + j] = __uint_as_float(v[j]);
    TMEM_LD16(tD3 + lane_off, v, 0); TMEM_LD16(tD3 + lane_off + 16, v, 16);
    TMEM_LD16(tD3 + lane_off + 32, v, 32); TMEM_LD16(tD3 + lane_off + 48, v, 48);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 64; ++j) D3[(size_t)tid * N1 + j] = __uint_as_float(v[j]);
  }
  // ---- test B: A operand from TMEM (S), B2 K-major in smem (two 32-wide K blocks)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t id = idesc_tf32(128, 32);
    for (int kb = 0; kb < 2; ++kb) {
      const uint64_t dB2 = desc_kmajor_sw128(smem_u32(sB2 + kb * 32 * 32));
      for (int ks = 0; ks < 4; ++ks) mma_ts(tD2, tS + kb * 32 + ks * 8, dB2 + 2 * ks, id, (kb | ks) > 0);
    }
    commit(&bar);
  }
  if (!wait_parity(&bar, 1)) { if (tid == 0) status[0] = 2; return; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t v[32];
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    TMEM_LD16(tD2 + lane_off, v, 0); TMEM_LD16(tD2 + lane_off + 16, v, 16);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D2[(size_t)tid * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
  if (tid == 0) status[0] = 0;
}

int main() {
  std::vector<float> A(128 * 32), B(N1 * 32), S(128 * 64), B2(32 * 64);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX; };
  for (auto& x : A) x = rnd() * 0.1f;
  for (auto& x : B) x = rnd();
  for (auto& x : S) x = (rnd() - 0.5f) * 8.0f;
  for (auto& x : B2) x = rnd();
  float *dA, *dB, *dS, *dB2, *dD1, *dD3, *dD2;
  int* dstat;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dS, S.size() * 4));
  CK(cudaMalloc(&dB2, B2.size() * 4)); CK(cudaMalloc(&dD1, 128 * N1 * 4)); CK(cudaMalloc(&dD3, 128 * N1 * 4));
  CK(cudaMalloc(&dD2, 128 * 32 * 4)); CK(cudaMalloc(&dstat, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dS, S.data(), S.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice));
  int st = -1;
  CK(cudaMemcpy(dstat, &st, 4, cudaMemcpyHostToDevice));
  const int smem = 49152 + 8192;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(dA, dB, dS, dB2, dD1, dD3, dD2, dstat);
  CK(cudaDeviceSynchronize());
  std::vector<float> D1(128 * N1), D3(128 * N1), D2(128 * 32);
  CK(cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost));
  printf("status %d\n", st);
  double e1 = 0, e3 = 0, e2 = 0, m1 = 0, m2 = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N1; ++j) {
      double r = 0;
      for (int k = 0; k < 32; ++k) r += (double)A[i * 32 + k] * B[j * 32 + k];
      e1 = fmax(e1, fabs(D1[i * N1 + j] - r)); e3 = fmax(e3, fabs(D3[i * N1 + j] - r)); m1 = fmax(m1, fabs(r));
    }
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 32; ++j) {
      double r = 0;
      for (int k = 0; k < 64; ++k) r += (double)S[i * 64 + k] * B2[j * 64 + k];
      e2 = fmax(e2, fabs(D2[i * 32 + j] - r)); m2 = fmax(m2, fabs(r));
    }
  printf("test A (SS tf32 x1): max abs err %.3e (rel to max %.3e)\n", e1, e1 / m1);
  printf("test A (SS tf32 x3): max abs err %.3e (rel to max %.3e)\n", e3, e3 / m1);
  printf("test B (TS tf32 x1): max abs err %.3e (rel to max %.3e)\n", e2, e2 / m2);
  printf("sample D1[0][0..3] = %g %g %g %g\n", D1[0], D1[1], D1[2], D1[3]);
  return 0;
}

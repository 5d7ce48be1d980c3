// Shared device helpers for the NBMF-MM sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace nbmf {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: a launcher keeps one bit per
// device and opts in again on every device it meets (a process-wide flag would leave cuda:1 without the opt-in after
// cuda:0 set it).  Safe to race: setting the attribute twice is harmless.
template <typename Kernel>
inline void ensure_dynamic_smem(Kernel kernel, int bytes, std::atomic<unsigned long long>& done_mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done_mask.load(std::memory_order_relaxed) & bit) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  done_mask.fetch_or(bit, std::memory_order_relaxed);
}

// workspace pointer of fit `blockIdx.z` of a batch (workspaces lie a fixed number of bytes apart)
template <typename T>
__device__ __forceinline__ T* batch_shift(T* p, size_t bytes) {
  return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) + bytes);
}


// ---------------------------------------------------------------- vector-of-2 arithmetic
// fp32 uses Blackwell's packed FFMA2 (PTX fma.rn.f32x2, sm_100+): one issue slot, two
// FMAs.  ptxas folds make2(s, s) operands into the scalar-broadcast form (Rn.F32), so
// broadcasting a ratio over a (k, k+1) accumulator pair costs no extra MOV.
template <typename Real> struct Vec2;
template <> struct Vec2<float>  { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

__device__ __forceinline__ float2 make2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ double2 make2(double a, double b) { return make_double2(a, b); }

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ double2 fma2(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y));
}

// ---------------------------------------------------------------- scalar math per precision
// fp32: MUFU.RCP / MUFU.LG2 (1 ulp-class); the log is accumulated in log2 units and scaled
// by ln 2 once per block.  fp64: IEEE division and log (parity mode, 1e-9 bar).
__device__ __forceinline__ float  rcp_(float x)  {   // bare MUFU.RCP: x >= eps here, never denormal
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// fp64: `1.0 / x` compiles to MUFU.RCP64H, a five-DFMA Newton chain AND a range test that branches to a slow path for
// operands outside the normal range; the branch cuts the row loop of the pass kernels into basic blocks that ptxas
// cannot interleave.  x = (Theta | 1 - Theta) + eps is always a normal number here, so a chain alone is enough, and three
// DFMAs of it: r0 = RCP64H(x) is good to ~2^-19 (it reads the high word of x), e = 1 - x r0, r = r0 (1 + e + e^2) leaves
// e^3 < 2^-56 -- within one ulp of the quotient (the parity bar is 1e-9); H pass 141 -> 117 ms, W pass 87 -> 77 ms at
// 10^5 x 10^5, K = 32 together with the hoisted loss-only flag (passes.cuh).
__device__ __forceinline__ double rcp_(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}
__device__ __forceinline__ float  div_(float a, float b)  { return a * rcp_(b); }
__device__ __forceinline__ double div_(double a, double b) { return a * rcp_(b); }   // dense V: b = x as above
__device__ __forceinline__ float  logu_(float x)  {   // bare MUFU.LG2 (log2 units)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double logu_(double x) { return log(x); }
template <typename Real> __device__ __forceinline__ double log_unit();
template <> __device__ __forceinline__ double log_unit<float>()  { return 0.693147180559945309417; }
template <> __device__ __forceinline__ double log_unit<double>() { return 1.0; }

// ---------------------------------------------------------------- cp.async (LDGSTS) staging
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- deterministic block sum (double)
// Fixed shuffle tree + fixed warp order: bit-reproducible run to run.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= NT/32 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += scratch[w];
  }
  return t;   // valid in thread 0
}

}  // namespace nbmf

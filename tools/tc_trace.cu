// Cycle trace of one CTA of the tensor H pass (development tool): arbitrary operand values, 148 CTAs,
// events of the MMA warp and of one SIMT warp per block group.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DTC_TRACE -Inbmf_mm_b200/csrc -o tools/bin/tc_trace tools/tc_trace.cu
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_passes_k32.cuh"
using namespace nbmf;
using namespace nbmf::k32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void fill(float* p, size_t n, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = scale * (float)((i * 2654435761u) % 1000u) * 1e-3f + 1e-3f;
}
__global__ void fillu(uint32_t* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = (uint32_t)(i * 2654435761u) & (uint32_t)((i * 40503u) >> 3);
}
__global__ void set_trace(long long* p) { g_tc_trace = p; }

int main(int argc, char** argv) {
  const int nblocks = argc > 1 ? atoi(argv[1]) : 512;       // 32-row blocks per CTA
  const int64_t m = 32LL * nblocks, n = 148 * 128, ldh = (n + 1023) / 1024 * 1024, mpad = (m + 127) / 128 * 128;
  float *H, *Wf, *CD; uint32_t* Pc; double* LL; int* done; long long* tr;
  CK(cudaMalloc(&H, 32 * ldh * 4)); CK(cudaMalloc(&Wf, mpad * 128 * 4)); CK(cudaMalloc(&CD, 2 * 32 * ldh * 4));
  CK(cudaMalloc(&Pc, ldh * mpad / 8)); CK(cudaMalloc(&LL, 148 * 8)); CK(cudaMalloc(&done, 4));
  CK(cudaMalloc(&tr, 4 * TC_TRACE_BLOCKS * 8 * 8));
  CK(cudaMemset(done, 0, 4)); CK(cudaMemset(tr, 0, 3 * TC_TRACE_BLOCKS * 8 * 8));
  fill<<<1024, 256>>>(H, 32 * ldh, 0.9f); fill<<<1024, 256>>>(Wf, mpad * 128, 0.03f); fillu<<<1024, 256>>>(Pc, ldh * mpad / 32);
  set_trace<<<1, 1>>>(tr);
  HTcArgs a; a.H = H; a.Wf = Wf; a.Pc = Pc; a.Mc = nullptr; a.m = m; a.n = n; a.ldh = ldh; a.nrb = mpad / 32; a.rows_per_split = mpad;
  a.CD = CD; a.LL = LL; a.eps = 1e-8f; a.done = done; a.compute_cd = 1; a.k = 32; a.colcnt = nullptr; a.flipcol = nullptr; a.flip_any = nullptr;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch_h_pass_tc<32>(a, 1, 0);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); launch_h_pass_tc<32>(a, 1, 0); cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("H pass: %d blocks/CTA, %.3f ms, %.1f cycles/block at 1.965 GHz, %.2f entries/clk/SM\n", nblocks, ms,
         ms * 1e-3 * 1.965e9 / nblocks, 4096.0 * nblocks / (ms * 1e-3 * 1.965e9));
  std::vector<long long> h(3 * TC_TRACE_BLOCKS * 8);
  CK(cudaMemcpy(h.data(), tr, h.size() * 8, cudaMemcpyDeviceToHost));
  auto ev = [&](int slot, int b, int e) { return h[(slot * TC_TRACE_BLOCKS + b) * 8 + e]; };
  const long long t0 = ev(0, 64, 0);
  printf("block | MMA1 warp: ready issued | MMA2 warp: s(b) issued | SIMT: top theta ld rfree_w rfree st2 arrive\n");
  for (int b = 64; b < 64 + 24 && b < nblocks; ++b) {
    printf("%4d | %6lld %6lld | %6lld %6lld |", b, ev(0, b, 0) - t0, ev(0, b, 1) - t0, ev(0, b, 2) - t0, ev(0, b, 3) - t0);
    const int g = b & 1;
    for (int e = 0; e < 7; ++e) printf(" %6lld", ev(1 + g, b, e) - t0);
    printf("\n");
  }
  return 0;
}

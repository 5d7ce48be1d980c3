// What does an mbarrier word look like?  (development experiment: can a plain ld.shared read the phase?)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../nbmf_mm_b200/csrc/tc_common.cuh"
using namespace nbmf::tc;
__global__ void k(unsigned long long* out) {
  __shared__ uint64_t bar;
  int n = 0;
  mbar_init(&bar, 3); mbar_fence_init();
  out[n++] = *(volatile uint64_t*)&bar;
  for (int ph = 0; ph < 3; ++ph)
    for (int a = 0; a < 3; ++a) { mbar_arrive(&bar); out[n++] = *(volatile uint64_t*)&bar; }
  mbar_expect_tx(&bar, 4096); out[n++] = *(volatile uint64_t*)&bar;
  // latency of a plain shared load of the word vs a try_wait probe of a completed phase
  long long t0 = clock64(); uint64_t w = *(volatile uint64_t*)&bar; long long t1 = clock64();
  bool ok = mbar_try(smem_u32(&bar), 1); long long t2 = clock64();
  out[n++] = (unsigned long long)(t1 - t0); out[n++] = (unsigned long long)(t2 - t1); out[n++] = ok + (w != 0);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 256); cudaMemset(d, 0, 256);
  k<<<1, 1>>>(d); cudaDeviceSynchronize();
  unsigned long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  printf("init(3)        %016llx\n", h[0]);
  for (int i = 1; i <= 9; ++i) printf("arrive #%d      %016llx%s\n", i, h[i], i % 3 == 0 ? "   <- phase completes" : "");
  printf("expect_tx 4096 %016llx\n", h[10]);
  printf("ld.shared latency %llu clk, try_wait(completed) latency %llu clk\n", h[11], h[12]);
  return 0;
}

// Host-side launchers of the tensor-core pass kernels (tc_passes.cuh).
#include "internal.h"
#include "tc_passes.cuh"

namespace nbmf {

void launch_w_pass_tensor(const WPassArgs& a, const void* Hf, const void* PM, int nsplit, cudaStream_t st) {
  WTcArgs t;
  t.W = (const float*)a.W; t.Hf = (const float*)Hf; t.PM = (const uint2*)PM;
  t.m = a.m; t.n = a.n; t.wpr = a.wpr;
  t.cols_per_split = a.cols_per_split;
  t.G = (float*)a.G; t.Q = (float*)a.Q; t.eps = (float)a.eps; t.done = a.done;
  launch_w_pass_tc(t, nsplit, st);
}

void launch_h_pass_tensor(const HPassArgs& a, const void* Wf, const uint32_t* Pc, const uint32_t* Mc, int64_t nrb,
                          int nsplit, cudaStream_t st) {
  HTcArgs t;
  t.H = (const float*)a.H; t.Wf = (const float*)Wf; t.Pc = Pc; t.Mc = Mc;
  t.m = a.m; t.n = a.n; t.ldh = a.ldh; t.nrb = nrb;
  t.rows_per_split = a.rows_per_split;
  t.CD = (float*)a.CD; t.LL = a.LL; t.eps = (float)a.eps; t.done = a.done; t.compute_cd = a.compute_cd;
  launch_h_pass_tc(t, nsplit, st);
}

}  // namespace nbmf

// Host-side launchers of the tensor-core pass kernels (tc_passes.cuh).
#include "internal.h"
#include "tc_passes.cuh"

namespace nbmf {

void launch_w_pass_tensor(const WPassArgs& a, const void* Wa, const void* Ha, const void* Hb, int nsplit,
                          cudaStream_t st) {
  WTcArgs t;
  t.f.Wa = (const float*)Wa; t.f.Ha = (const float*)Ha; t.f.Hb = (const float*)Hb; t.f.Wb = nullptr;
  t.P = a.P; t.M = a.M; t.m = a.m; t.n = a.n; t.wpr = a.wpr;
  t.cols_per_split = a.cols_per_split;
  t.G = (float*)a.G; t.Q = (float*)a.Q; t.eps = (float)a.eps; t.done = a.done;
  launch_w_pass_tc(t, nsplit, st);
}

void launch_h_pass_tensor(const HPassArgs& a, const void* Ha, const void* Wa, const void* Wb, const uint32_t* Pt,
                          int64_t wpr_t, int nsplit, cudaStream_t st) {
  HTcArgs t;
  t.f.Ha = (const float*)Ha; t.f.Wa = (const float*)Wa; t.f.Wb = (const float*)Wb; t.f.Hb = nullptr;
  t.Pt = Pt; t.m = a.m; t.n = a.n; t.ldh = a.ldh; t.wpr_t = wpr_t;
  t.rows_per_split = a.rows_per_split;
  t.CD = (float*)a.CD; t.LL = a.LL; t.eps = (float)a.eps; t.done = a.done; t.compute_cd = a.compute_cd;
  launch_h_pass_tc(t, nsplit, st);
}

}  // namespace nbmf

// Host-side launchers of the tensor-core pass kernels: K <= 32 (tc_passes_k32.cuh: two full-block pipelines of eight
// SIMT warps, the fastest structure TMEM has room for at K <= 32) and 32 < K <= 64 (tc_passes_k64.cuh: three half-block
// pipelines of four warps and one accumulator set -- what fits beside a 128-column resident operand).
#include <stdlib.h>

#include "internal.h"
#include "tc_passes_k64.cuh"
#include "tc_passes_k32.cuh"

namespace nbmf {

void launch_w_pass_tensor(const WPassArgs& a, const void* Hf, const void* PM, int kb, int nsplit, cudaStream_t st) {
  WTcArgs t;
  t.W = (const float*)a.W; t.Hf = (const float*)Hf; t.PM = (const uint2*)PM;
  t.m = a.m; t.n = a.n; t.wpr = a.wpr;
  t.cols_per_split = a.cols_per_split;
  t.G = (float*)a.G; t.Q = (float*)a.Q; t.eps = (float)a.eps; t.done = a.done;
  t.batch_stride = a.batch_n > 1 ? a.batch_stride : 0;
  if (kb == 16) k32::launch_w_pass_tc<16>(t, nsplit, a.batch_n, st);
  else if (kb == 32) k32::launch_w_pass_tc<32>(t, nsplit, a.batch_n, st);
  else k64::launch_w_pass_tc_kb<64>(t, nsplit, a.batch_n, st);
}

void launch_h_pass_tensor(const HPassArgs& a, const void* Wf, const uint32_t* Pc, const uint32_t* Mc, int64_t nrb,
                          int kb, int k, const uint32_t* colcnt, uint32_t* flipcol, int* flip_any, int nsplit, cudaStream_t st) {
  HTcArgs t;
  t.H = (const float*)a.H; t.Wf = (const float*)Wf; t.Pc = Pc; t.Mc = Mc;
  t.m = a.m; t.n = a.n; t.ldh = a.ldh; t.nrb = nrb;
  t.rows_per_split = a.rows_per_split;
  t.CD = (float*)a.CD; t.LL = a.LL; t.eps = (float)a.eps; t.done = a.done; t.compute_cd = a.compute_cd;
  t.k = k; t.colcnt = colcnt; t.flipcol = flipcol; t.flip_any = flip_any;
  t.batch_stride = a.batch_n > 1 ? a.batch_stride : 0;
  if (kb == 16) k32::launch_h_pass_tc<16>(t, nsplit, a.batch_n, st);
  else if (kb == 32) k32::launch_h_pass_tc<32>(t, nsplit, a.batch_n, st);
  else k64::launch_h_pass_tc_kb<64>(t, nsplit, a.batch_n, st);
}

}  // namespace nbmf

// Argument block of the fused small-fit kernel (fused_small.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "internal.h"

namespace nbmf {

struct FusedArgs {
  void* W;                   // [m][kp]
  void* H;                   // [kp][ldh]
  void* Ht;                  // [ldh][kp]   kept in step for the entry points that use the regular W pass
  const uint32_t* P;         // [m][wpr] bit planes (shared by the fits of a batch)
  const uint32_t* M;         // observation mask or NULL
  int64_t m, n, ldh, wpr;
  int k, kp, strict, projection;
  int rows_per_warp;         // R: a CTA unit of the H phase is 8 R rows x one 32-column word
  int nsuper;                // row blocks: ceil(m / 8R)
  int nwords;                // ceil(n / 32)
  void* CDpart;              // [nsuper][2][kp][32 nwords]  partial C | D of the row blocks
  double* LLpart;            // [nwords * nsuper]           partial log-likelihoods
  double* prior_part;        // [n_prior][2]                partial sums of log(H + eps), log((1 - H) + eps)
  double* prior_part2;       // second buffer of the same: the fused kernel alternates them by iteration parity
  int n_prior;
  int wsplit;                // W phase: warps that share a row (1, 2, 4 or 8; from the shape only)
  FitState* state;
  double* history;
  const void* rowcount;      // Duchi with a mask: observed entries per row, else NULL
  double eps, n_obs, tol;
  int max_iter;
  unsigned* bar;             // grid barrier counter, 0 at launch
  int n_passes;              // H passes this launch may run (one per iteration + the loss-only pass after the last)
  int h_in_smem;             // the W phase stages H (k x 32 nwords) in shared memory
  int64_t batch_stride;      // bytes between the workspaces of the fits of a batch (blockIdx.y)
  unsigned long long* trace; // development hook (env NBMF_FUSED_TRACE): 8 globaltimer stamps per iteration of CTA 0, or NULL
};

int fused_kp(int k);
size_t fused_smem_bytes(int dtype, int k, int rows_per_warp, int nwords, int h_in_smem);
int fused_max_blocks(int dtype, const FusedArgs& a, size_t smem);
int launch_fused_fit(int dtype, const FusedArgs& a, int grid_x, int batch_n, size_t smem, cudaStream_t st);

}  // namespace nbmf

// Argument blocks of the tensor-core pass kernels (tc_passes.cuh).  KT = 32 (K <= 32) or 64 (K <= 64) is the K extent of
// the formatted operands and the padded K of W, H, C, D and G.
#pragma once
#include <stdint.h>

namespace nbmf {

struct HTcArgs {
  const float* H;            // [KT][ldh] k-major factor (pad rows/columns 0.5): source of the resident A tile
  const float* Wf;           // [mpad/32][4][KT/32][1024]  W rows hi | corr | W^T hi | corr   (format_factors.cu)
  const uint32_t* Pc;        // [ldh/128][nrb][128] bit r of word (jt, rb, jj) = P[32 rb + r][128 jt + jj]
  const uint32_t* Mc;        // same tiling of the observation mask (strict mask semantics only)
  int64_t m, n, ldh, nrb;
  int64_t rows_per_split;    // multiple of 32
  float* CD;                 // [nsplit][2][KT][ldh]
  double* LL;                // [nsplit * gridDim.x]
  float eps;
  const int* done;
  int compute_cd;
  int k;                     // n_components (rows of H that are real)
  const uint32_t* colcnt;    // [ldh] ones per column of P over this context's rows, or NULL: picks the plane (ones / zeros)
                             // that is accumulated directly, see h_pass_tc_kernel (K <= 64 kernels decide in the kernel)
  const uint32_t* flipcol;   // [ldh] K <= 32 kernels: the decision per column (0 / ~0), made by flip_cols_kernel per H pass
  const int* flip_any;       // 1 if any column of flipcol is set: selects the kernel instantiation that does the work
  // batched small fits: gridDim.z fits whose workspaces lie batch_stride bytes apart advance with one launch; EVERY pointer
  // above lives in the fit's workspace (factors, formatted operands, re-tiled planes, partials, flags) and is shifted by
  // blockIdx.z * batch_stride
  int64_t batch_stride = 0;
};

struct WTcArgs {
  const float* W;            // [m][KT] row-major factor: source of the resident A tile
  const float* Hf;           // [ldh/64][4][KT/32][2048]  Ht rows hi | corr | H (2 K-blocks) hi | corr
  const uint2* PM;           // [mpad/128][wpr][128] {P word, observed word} of row 128 it + ii, columns 32 cw..
  int64_t m, n, wpr;
  int64_t cols_per_split;    // multiple of 64
  float* G;                  // [nsplit][m][KT]
  float* Q;                  // [nsplit][m]
  float eps;
  const int* done;
  int64_t batch_stride = 0;  // see HTcArgs: W, Hf, PM, G, Q, done are shifted per fit
};

}  // namespace nbmf

// Kernel argument blocks shared by the pass kernels and the C-ABI dispatcher.
#pragma once
#include <stdint.h>

namespace nbmf {

// Device-resident layout (all leading dimensions padded so that tiles never leave the
// allocation; see DESIGN.md "data layout"):
//   W   [m][KP]      Real  row-major, KP = K rounded up to the variant's K tile, pad = 0
//   H   [KP][ldh]    Real  k-major,   ldh = round_up(n, 1024), pad columns/rows = 0.5
//   Ht  [ldh][KP]    Real  transposed copy of H for the W pass (one row per column j)
//   P   [m][wpr]     u32   bit j of row i = V[i][j] & mask[i][j]   (wpr = ldh / 32)
//   M   [m][wpr]     u32   observation mask bits, or NULL when everything is observed
//   Vm  [m][ldv]     Real  V * mask for probabilistic V (dense mode), ldv = ldh, pad = 0
struct HPassArgs {
  const void* W;
  const void* H;
  const uint32_t* P;
  const uint32_t* M;
  const void* Vm;
  int64_t ldv;
  int64_t m, n, ldh, wpr;
  int64_t rows_per_split;
  void* CD;          // [nsplit][2][KP][ldh] partial numerators (C) and denominators (D)
  double* LL;        // [nsplit * gridDim.x] partial log-likelihoods (natural log units)
  double eps;
  const int* done;   // device flag: 1 = converged, every kernel becomes a no-op
  int compute_cd;    // 0 = loss-only pass
  // batched small fits (SIMT engine): gridDim.z fits whose workspaces lie batch_stride bytes apart advance with one
  // launch; every workspace pointer above (W, H, CD, LL, done) is shifted by blockIdx.z * batch_stride, the data
  // planes are shared
  int batch_n = 1;
  int64_t batch_stride = 0;
};

struct WPassArgs {
  const void* W;
  const void* Ht;
  const uint32_t* P;
  const uint32_t* M;
  const void* Vm;
  const void* Wm = nullptr;  // dense mode, weighted mask: the mask VALUES [m][ldv] Real (pad 0); NULL = 0/1 mask (bit plane M)
  int64_t ldv;
  int64_t m, n, ldh, wpr;
  int64_t cols_per_split;   // multiple of 128
  void* G;           // [nsplit][m][KP] partial sum_j H[k][j] * (p - q)
  void* Q;           // [nsplit][m]     partial sum_j q
  double eps;
  const int* done;
  int batch_n = 1;           // see HPassArgs: W, Ht, G, Q, done are shifted per fit
  int64_t batch_stride = 0;
};

struct PassLaunch {
  void (*h_launch)(const HPassArgs&, int nsplit, cudaStream_t);
  void (*w_launch)(const WPassArgs&, int nsplit, cudaStream_t);
  int h_bn;      // columns per H-pass CTA
  int w_bmr;     // rows per W-pass CTA
  int kp;        // padded K of this variant
};

}  // namespace nbmf

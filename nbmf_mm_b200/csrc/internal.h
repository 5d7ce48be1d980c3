// Internal C++ interface between the C-ABI (capi.cu), the pass instantiations and the
// small epilogue / data-layer kernels.  Nothing here is exported.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "args.h"

namespace nbmf {

// ---- per-fit device state (one struct in device memory; read by every kernel)
struct FitState {
  int done;            // 1 = stop (converged or max_iter reached); kernels become no-ops
  int it;              // number of H passes finalised so far == MM iterations started
  int n_hist;          // number of losses recorded (== n_iter_ when done)
  int converged;       // 1 = stopped by the tolerance rule
  double prev_loss;    // loss of the previous iteration (inf before the first)
  double prior_a;      // sum log(H + eps) of the current H
  double prior_b;      // sum log((1 - H) + eps) of the current H
  double alpha, beta;  // Beta prior of this fit (_solver.py:35-36): per fit, so that the fits of a batch (gridDim.z) may differ
};

// ---- loss + stop rule of one iteration (finalize_body in misc_kernels.cu); state == nullptr: not requested
struct FinalizeArgs {
  FitState* state;
  const double* prior_part;
  int n_prior_part;
  double alpha, beta, n_obs, tol;
  int max_iter;
  double* history;
};

// ---- variant lookup: dtype 0 = f32, 1 = f64; returns false if K is unsupported
bool lookup_pass(int dtype, int dense, int strict, int k, PassLaunch* out);

// ---- epilogues and small reductions (misc_kernels.cu); dtype as above
void launch_init_factors(int dtype, const void* W_in, const void* H_in, int64_t m, int64_t n, int k, int kp,
                         int64_t ldh, void* W, void* H, void* Ht, int normalize_w, cudaStream_t st);
void launch_export_factors(int dtype, const void* W, const void* H, int64_t m, int64_t n, int k, int kp,
                           int64_t ldh, void* W_out, void* H_out, cudaStream_t st);
void launch_h_reduce(int dtype, const void* CDpart, int nsplit, int64_t count, void* CDsum,
                     const double* LLpart, int64_t n_ll, double* LLsum, const FitState* state, const FinalizeArgs& fin,
                     cudaStream_t st, int batch_n = 1, int64_t batch_stride = 0);
void launch_finalize(const FinalizeArgs& f, const double* LLsum, cudaStream_t st);
int  h_epilogue_blocks(int64_t n, int kp);
void launch_h_epilogue(int dtype, const void* CDsum, int64_t n, int k, int kp, int64_t ldh, double alpha,
                       double beta, double eps, void* H, void* Ht, double* prior_part, const FitState* state,
                       cudaStream_t st, int batch_n = 1, int64_t batch_stride = 0);
void launch_prior_sums(int dtype, const void* H, int64_t n, int k, int kp, int64_t ldh, double eps,
                       double* prior_part, cudaStream_t st);
void launch_w_epilogue(int dtype, const void* Gpart, const void* Qpart, int nsplit, int64_t m, int64_t n,
                       int k, int kp, int projection, const void* rowcount, void* W, const FitState* state,
                       cudaStream_t st, int batch_n = 1, int64_t batch_stride = 0);
void launch_simplex_deviation(int dtype, const void* W, int64_t m, int k, int kp, unsigned long long* out, cudaStream_t st);
void launch_export_f64(int dtype, const void* W, const void* H, int64_t m, int64_t n, int k, int kp, int64_t ldh,
                       int normalize_w, double* W_out, double* H_out, cudaStream_t st);
void launch_clip_rows(int dtype, void* W, int64_t m, int k, int kp, double lo, double hi, cudaStream_t st);

// ---- tensor-core engine (K <= 64, fp32, bit-packed V): operand formatting + pass launchers
void launch_format_w(const void* W, int64_t m, int64_t mpad, int kt, void* Wf, const FitState* state, cudaStream_t st,
                     int batch_n = 1, int64_t batch_stride = 0);
void launch_format_h(const void* H, int64_t ldh, int kt, float theta_bias, void* Hf, const FitState* state, cudaStream_t st,
                     int batch_n = 1, int64_t batch_stride = 0);
void launch_colcount(const uint32_t* Pc, int64_t nrb, int64_t rb0, int64_t rb1, int64_t ldh, uint32_t* colcnt, cudaStream_t st);
void launch_tile_planes(const uint32_t* P, const uint32_t* M, int64_t m, int64_t n, int64_t wpr, int64_t mpad,
                        int64_t row0, int64_t row1, uint32_t* Pc, uint32_t* Mc, void* PM, cudaStream_t st);
// kb = 16 | 32 | 64: K extent the MMAs cover (TcCfg<KB> in tc_passes.cuh); k = n_components
void launch_w_pass_tensor(const WPassArgs& a, const void* Hf, const void* PM, int kb, int nsplit, cudaStream_t st);
void launch_h_pass_tensor(const HPassArgs& a, const void* Wf, const uint32_t* Pc, const uint32_t* Mc, int64_t nrb,
                          int kb, int k, const uint32_t* colcnt, uint32_t* flipcol, int* flip_any, int nsplit, cudaStream_t st);

// ---- data layer
void launch_pack_bits(int in_dtype, const void* X, int64_t ldx, const void* mask, int mask_dtype, int64_t ldm,
                      int64_t m, int64_t n, int64_t wpr, uint32_t* P, uint32_t* M, int* flags, cudaStream_t st);
void launch_pack_dense(int in_dtype, const void* X, int64_t ldx, const void* mask, int mask_dtype, int64_t ldm,
                       int64_t m, int64_t n, int out_dtype, int64_t ldv, void* Vm, cudaStream_t st);
void launch_pack_csr(const int64_t* indptr, const int32_t* indices, const void* data, int data_dtype, int64_t m,
                     int64_t n, int64_t wpr, uint32_t* P, int* flags, cudaStream_t st);
void launch_reconstruct(int dtype, const void* W, const void* H, int64_t m, int64_t n, int k, void* out, cudaStream_t st);
void launch_transpose_bits(const uint32_t* src, int64_t m, int64_t n, int64_t wpr_src, uint32_t* dst,
                           int64_t wpr_dst, cudaStream_t st);
void launch_rowcount(int dtype, const uint32_t* M, int64_t m, int64_t n, int64_t wpr, void* out, cudaStream_t st);
void launch_popcount(const uint32_t* B, int64_t m, int64_t wpr, unsigned long long* out, cudaStream_t st);
// streamed ingestion of host planes: P &= M in place and *count += popcount(M) over `words` words
void launch_and_count(uint32_t* P, const uint32_t* M, int64_t words, unsigned long long* count, cudaStream_t st);
void launch_synth_bits(uint64_t seed, int64_t row0, int64_t m, int64_t n, int64_t wpr, const float* Wstar,
                       const float* Hstar, int kstar, float obs_frac, uint32_t* P, uint32_t* M, cudaStream_t st);
double run_fma_peak(int dtype, int iters, cudaStream_t st, float* scratch);

}  // namespace nbmf

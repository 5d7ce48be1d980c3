/* nbmf_b200.h -- C-ABI of libnbmf_b200.so: the B200 (sm_100a) implementation of the
 * NBMF-MM fit loop (mean-parameterised Bernoulli NMF, Magron & Fevotte 2022).
 *
 * This is the drop-in boundary for ONE path of siddC/nbmf_mm: what sits below
 * `NBMFMM.fit` (reference src/nbmf_mm/_base.py:98-111), i.e. `nbmf_mm_solver`
 * (src/nbmf_mm/_solver.py:61-216), its one-step function `nbmf_mm_update_beta_dir`
 * (_solver.py:5-59) and the fixed-H W-solver inlined in `NBMFMM.transform`
 * (_base.py:178-198).  The reference has no FFI of its own (pure NumPy); each entry point
 * below names the reference code it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer named *_dev is a DEVICE pointer on the current CUDA device; the caller
 *    owns all memory (the library never allocates device memory: the caller sizes a
 *    workspace with nbmf_workspace_bytes and passes it to nbmf_create);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, only the calls
 *    documented as synchronising wait for it;
 *  - matrices are in the solver's INTERNAL orientation (dir-beta callers pass the packed
 *    transpose, as _solver.py:113-123 does): V is m x n, W is m x k row-major, H is k x n
 *    row-major; `dtype` 0 = float32, 1 = float64 selects the arithmetic type of the path;
 *  - bit planes are uint32 words, bit (j % 32) of word (j / 32) of row i, rows padded to
 *    nbmf_words_per_row(n) words (a multiple of 32 words = 1024 columns), padding bits zero;
 *  - return value 0 = ok, negative = error; nbmf_last_error() gives the message.  There is
 *    no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef NBMF_B200_H
#define NBMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBMF_OK 0
#define NBMF_ERR_ARG (-1)
#define NBMF_ERR_CUDA (-2)
#define NBMF_ERR_UNSUPPORTED (-3)
#define NBMF_ERR_NCCL (-4)

#define NBMF_F32 0
#define NBMF_F64 1
#define NBMF_U8 2
#define NBMF_F16 3      /* storage only: dense V*mask of the NBMF_V_DENSE_F16 layout */

#define NBMF_V_BITS 0   /* binary V: planes P = V & mask and M = mask, 1 bit per entry each */
#define NBMF_V_DENSE 1  /* probabilistic V in [0,1]: dense V*mask in `dtype` + mask bit plane */
#define NBMF_V_DENSE_F16 2 /* same, V*mask stored as fp16 (half the HBM bytes; float32 arithmetic only) */

#define NBMF_MASK_REFERENCE 0 /* H step / loss treat unobserved entries as observed zeros (_solver.py:43,153-154) */
#define NBMF_MASK_STRICT 1    /* only observed entries contribute (README / paper); unpinned */

#define NBMF_PROJ_NORMALIZE 0 /* multiplicative step, /n, L1 renormalisation (_solver.py:53-57) */
#define NBMF_PROJ_DUCHI 1     /* multiplicative step / n_obs(row), Euclidean simplex projection; unpinned */

#define NBMF_ENGINE_AUTO 0   /* a single small fit: FUSED; else TENSOR when eligible and m >= 512, n >= 128; else SIMT */
#define NBMF_ENGINE_SIMT 1   /* packed-FFMA2 CUDA-core kernels: every dtype / K <= 128 / layout */
#define NBMF_ENGINE_TENSOR 2 /* tcgen05 + TMEM kernels, TF32 + bf16 split precision: float32, bit-packed V, K <= 64, eps >= 1e-9 */
#define NBMF_ENGINE_FUSED 3  /* SIMT arithmetic, but the fit loop (nbmf_fit / nbmf_fit_enqueue) runs whole iterations inside one
                              * persistent cooperative kernel: bit-packed V, K <= 32, single GPU.  AUTO picks it while
                              * m * n * padded K <= 1.6e7 (fp64) / 1.2e7 (fp32): a fit that small is bound by launch latency.
                              * The half-step / transform entry points of such a context use the SIMT kernels. */

typedef struct nbmf_ctx nbmf_ctx;

/* Problem description; replaces the keyword arguments of nbmf_mm_solver (_solver.py:61-75). */
typedef struct nbmf_config {
  int64_t m;             /* local rows of this shard (internal orientation) */
  int64_t n;             /* columns */
  int32_t k;             /* n_components, 1..128 (tensor engine: <= 64, persistent small-fit kernel: <= 32) */
  int32_t dtype;         /* NBMF_F32 | NBMF_F64 */
  int32_t vkind;         /* NBMF_V_BITS | NBMF_V_DENSE | NBMF_V_DENSE_F16 */
  int32_t mask_semantics;
  int32_t projection;
  int32_t has_mask;      /* 0: everything observed (mask=None) */
  double alpha, beta;    /* Beta prior of H (_solver.py:35-36) */
  double eps;            /* 1e-8 in the reference */
  double n_obs;          /* loss denominator: Y.size or count_nonzero(mask) over ALL shards (_solver.py:151,155) */
  int32_t max_iter_cap;  /* capacity of the on-device loss history */
  int32_t engine;        /* NBMF_ENGINE_AUTO | _SIMT | _TENSOR | _FUSED (env NBMF_ENGINE=auto|simt|tensor|fused overrides) */
  int32_t batch_hint;    /* 0 / 1: plan the launches for this fit alone.  n > 1: the context will be one of n fits that
                          * advance together (nbmf_batch_bind): the batch fills the SMs, so a fit is cut into fewer, larger
                          * row / column splits (config 5: 64 restarts 25 % faster).  Changes the summation order of the
                          * split partials, i.e. results agree with an unbatched fit to rounding, not bit for bit. */
} nbmf_config;

int nbmf_version(void);
const char* nbmf_last_error(void);
int64_t nbmf_words_per_row(int64_t n);          /* uint32 words per bit-plane row (multiple of 32) */
int64_t nbmf_padded_cols(int64_t n);            /* leading dimension of dense V*mask rows (= 32 * words) */

/* ---- data layer (replaces the per-iteration Y*mask / transposes of _solver.py:21-32) ---- */
/* dense X (x_dtype F32|F64|U8, leading dim ldx) [+ mask] -> bit planes P (= X!=0 & mask) and M (nullable) */
int nbmf_pack_bits(const void* x_dev, int x_dtype, int64_t ldx, const void* mask_dev, int mask_dtype, int64_t ldm,
                   int64_t m, int64_t n, uint32_t* p_bits_dev, uint32_t* m_bits_dev, void* stream);
/* The same for a block of rows, with the value checks of the reference front end done on the device instead of by
 * NumPy passes over X and the mask (check of _base.py:90-91 "X must be binary", binary mask): *flags_dev (int32,
 * caller-zeroed, accumulated over calls) gets bit 0 = X holds a value that is not 0 or 1 (probabilistic V), bit 1 = X
 * holds a value outside [0, 1] or a NaN, bit 2 = the mask holds a value that is not 0 or 1.  No synchronisation. */
int nbmf_pack_bits_checked(const void* x_dev, int x_dtype, int64_t ldx, const void* mask_dev, int mask_dtype, int64_t ldm,
                           int64_t m, int64_t n, uint32_t* p_bits_dev, uint32_t* m_bits_dev, int32_t* flags_dev, void* stream);
/* dense X [+ mask] -> V*mask in out_dtype with leading dimension nbmf_padded_cols(n), zero padded */
int nbmf_pack_dense(const void* x_dev, int x_dtype, int64_t ldx, const void* mask_dev, int mask_dtype, int64_t ldm,
                    int64_t m, int64_t n, int out_dtype, void* vm_dev, void* stream);
/* CSR matrix (int64 indptr[m+1], int32 indices, optional f32/f64 data; all device pointers) -> bit plane, no dense
 * M x N intermediate (the reference densifies sparse X and mask: _base.py:83-87, _solver.py:106-107).  Explicit
 * zeros stay zero bits.  *flags_host: bit 0 = a stored value is not 0 or 1, bit 1 = a stored value is outside [0, 1].
 * Synchronises the stream. */
int nbmf_pack_csr(const int64_t* indptr_dev, const int32_t* indices_dev, const void* data_dev, int data_dtype, int64_t m,
                  int64_t n, uint32_t* p_bits_dev, int32_t* flags_dev, int32_t* flags_host, void* stream);
/* inverse_transform (_base.py:201-210): out (m x n, row-major, dtype) = clip(W (m x k) @ H (k x n), 0, 1) */
int nbmf_reconstruct(int dtype, const void* w_dev, const void* h_dev, int64_t m, int64_t n, int32_t k, void* out_dev, void* stream);
/* bit-plane transpose (m x n bits -> n x m bits): dir-beta runs as beta-dir on V^T (_solver.py:113-123) */
int nbmf_transpose_bits(const uint32_t* src_dev, int64_t m, int64_t n, uint32_t* dst_dev, void* stream);
/* number of set bits of a plane (count_nonzero(mask), _solver.py:155); synchronises the stream */
int nbmf_popcount_bits(const uint32_t* bits_dev, int64_t m, int64_t n, uint64_t* scratch_dev, uint64_t* count_host,
                       void* stream);
/* counter-based synthetic generator keyed on (seed, global row, column): V ~ Bernoulli(W*H*) with
 * Dirichlet(1) rows W* (derived from the key) and the given H* (kstar x n, float32), mask ~ Bernoulli(obs_frac) */
int nbmf_synth_bits(uint64_t seed, int64_t row0, int64_t m, int64_t n, const float* hstar_dev, int32_t kstar,
                    float obs_frac, uint32_t* p_bits_dev, uint32_t* m_bits_dev, void* stream);

/* ---- fit context ---- */
int64_t nbmf_workspace_bytes(const nbmf_config* cfg);
int nbmf_create(const nbmf_config* cfg, void* workspace_dev, int64_t workspace_bytes, void* stream, nbmf_ctx** out);
int nbmf_destroy(nbmf_ctx* ctx);
/* borrow the data planes (must outlive the context) */
int nbmf_set_data_bits(nbmf_ctx* ctx, const uint32_t* p_bits_dev, const uint32_t* m_bits_dev);
/* Streamed ingestion of HOST planes (the check_array / densify / `Y * mask` front end of _base.py:83-93 and
 * _solver.py:106-110, for inputs that are already bit-packed on the host).  The caller copies rows [row0, row1) of the
 * V plane into p_bits_dev (and of the mask into m_bits_dev) on its own copy stream, orders the context's stream after
 * the copy (event) and calls nbmf_ingest_bits_rows: P &= M in place, the mask count (count_nonzero(mask),
 * _solver.py:155) and the re-tiling for the tensor engine then run while later chunks are still crossing PCIe.  Chunks
 * are consecutive, start on multiples of 128 rows and the last one ends at m.  nbmf_ingest_bits_end synchronises the
 * stream and returns the mask count (m * n without a mask); nbmf_set_n_obs installs the normaliser of the objective
 * (the count itself, or its sum over the row shards of a multi-GPU fit). */
/* 1 while the context still reads the caller's row-major planes, 0 once it does not: the tensor engine works on its own
 * re-tiled copies (and takes the per-row observed counts of the Duchi projection when the planes are handed over), so the
 * caller may free or reuse the planes after nbmf_set_data_bits / nbmf_ingest_bits_end -- stream order makes a free or an
 * overwrite enqueued on the same stream afterwards safe.  At config 4 that is 25 GB of the 62.5 GB a fit would hold. */
int nbmf_planes_in_use(nbmf_ctx* ctx);
int nbmf_ingest_bits_begin(nbmf_ctx* ctx, uint32_t* p_bits_dev, const uint32_t* m_bits_dev);
int nbmf_ingest_bits_rows(nbmf_ctx* ctx, int64_t row0, int64_t row1);
int nbmf_ingest_bits_end(nbmf_ctx* ctx, double* mask_count_host);
int nbmf_set_n_obs(nbmf_ctx* ctx, double n_obs);
int nbmf_set_data_dense(nbmf_ctx* ctx, const void* vm_dev, const uint32_t* m_bits_dev);
/* Weighted observation mask (values other than 0 / 1): the reference multiplies by the mask VALUES (Y * mask for the H
 * half-step and the loss, (1 - Y).T * mask.T for the W half-step: _solver.py:30-32; count_nonzero(mask) normalises the
 * loss, :155).  V * mask is what nbmf_set_data_dense takes (NBMF_V_DENSE layout in cfg.dtype, also for a binary V), its bit
 * plane is then (mask != 0); wm_dev holds the mask values in the same dense layout and dtype ((1 - V) * mask = mask -
 * V * mask).  Call after nbmf_set_data_dense; NULL goes back to a 0/1 mask.  Reference mask semantics only. */
int nbmf_set_mask_weights(nbmf_ctx* ctx, const void* wm_dev);
/* W_init (m x k) / H_init (k x n) in cfg.dtype; normalize_w != 0 divides every W row by its sum
 * (_solver.py:132-136).  Either pointer may be NULL to keep the current factor.  Resets the loop state. */
int nbmf_set_factors(nbmf_ctx* ctx, const void* w_dev, const void* h_dev, int normalize_w);
int nbmf_get_factors(nbmf_ctx* ctx, void* w_dev, void* h_dev);
/* tail of the reference solver on the device (_solver.py:192-213): worst |rowsum(W) - 1| in fp64 (NaN if a row sum is
 * not finite; synchronises), and the export of W (m x k) / H (k x n) as fp64 with W's rows optionally divided by their
 * sums (rows with sum <= 1e-12 are left alone).  The caller applies the reference's rule: normalise iff dev > 1e-9. */
int nbmf_simplex_deviation(nbmf_ctx* ctx, double* dev_host);
int nbmf_get_factors_f64(nbmf_ctx* ctx, double* w_dev, double* h_dev, int normalize_w);

/* ---- single steps (replace nbmf_mm_update_beta_dir, _solver.py:5-59) ---- */
int nbmf_h_half_step(nbmf_ctx* ctx);            /* H <- H' (_solver.py:39-47) */
int nbmf_w_half_step(nbmf_ctx* ctx);            /* W <- W' with the current H (_solver.py:50-57) */
/* MAP objective of the current factors (_solver.py:148-162); synchronises the stream */
int nbmf_objective(nbmf_ctx* ctx, double* loss_host);
/* Per-CTA partial sums of the log-likelihood (the `np.sum(log_lik)` of _solver.py:150-161 before the sum) left by the
 * most recent H pass: entry [s * col_blocks + b] covers the rows of row split s and the columns of column block b
 * (nbmf_plan_info / nbmf_variant_info give the block width).  Lets a caller check the fused NLL of one column block
 * against a host computation on a problem too large to check whole.  Synchronises the stream. */
int nbmf_loglik_partials(nbmf_ctx* ctx, double* out_host, int64_t capacity, int32_t* col_blocks_host, int32_t* row_splits_host);

/* ---- the fit loop (replaces the loop of nbmf_mm_solver, _solver.py:143-175) ----
 * Runs up to max_iter MM iterations entirely on the stream; the loss of every iteration, the
 * relative-change stop rule and n_iter are evaluated on the device.  Synchronises at the end and
 * copies the loss history (n_iter doubles) and n_iter to the host. */
int nbmf_fit(nbmf_ctx* ctx, int32_t max_iter, double tol, double* history_host, int32_t* n_iter_host,
             int32_t* converged_host);
/* same loop, asynchronous: enqueue `n_iters` more iterations (plus the trailing loss pass when the
 * budget max_iter is reached) and return without waiting */
int nbmf_fit_begin(nbmf_ctx* ctx, int32_t max_iter, double tol);
int nbmf_fit_enqueue(nbmf_ctx* ctx, int32_t n_iters);
/* non-blocking unless wait != 0: fetch (done, n_iter) of the work enqueued so far */
int nbmf_fit_poll(nbmf_ctx* ctx, int wait, int32_t* done_host, int32_t* n_iter_host);
/* Batched small fits (n_init restarts, alpha / beta grids over one data set): n contexts of identical configuration
 * (alpha and beta may differ from context to context) and data planes, workspaces at a uniform byte stride inside one allocation (the leader's
 * first), each with its own factors (nbmf_set_factors) and nbmf_fit_begin.  After nbmf_batch_bind(leader, n, stride)
 * the leader's nbmf_fit_enqueue advances all n fits with one launch per kernel (the Python loop of solver calls in
 * examples/reproduce_magron2022.py:87-117 / README.md:133,144 becomes 5 launches per iteration for the whole batch);
 * every fit keeps its own loss history, stop rule and n_iter on the device.  nbmf_batch_poll synchronises the stream
 * and reports whether every fit has stopped and how many losses each has recorded; results are then read from each
 * context as usual.  SIMT engine, single GPU. */
int nbmf_batch_bind(nbmf_ctx* leader, int32_t n, int64_t stride_bytes);
int nbmf_batch_poll(nbmf_ctx* leader, int32_t* all_done_host, int32_t* n_iter_host);
/* the tail of the solver (_solver.py:178-213) for every fit of the batch with one synchronisation: loss histories
 * (hist_stride doubles per fit), converged flags and max |row sum of W - 1| (NaN if a row sum is not finite) */
int nbmf_batch_tail(nbmf_ctx* leader, double* history_host, int32_t hist_stride, int32_t* converged_host,
                    double* deviation_host);
int nbmf_fit_history(nbmf_ctx* ctx, double* history_host, int32_t count, int32_t* converged_host);

/* ---- transform (replaces the 50 fixed-H W steps of NBMFMM.transform, _base.py:178-198) ----
 * n_steps W half-steps with the current H from the current W, then clip to [1e-8, 1] and row-normalise. */
int nbmf_transform(nbmf_ctx* ctx, int32_t n_steps);

/* ---- multi-GPU: rows are sharded, the K x N partials of the H step are summed with one
 * ncclAllReduce per iteration (no reference analogue) ---- */
int nbmf_comm_unique_id(void* id128_host);                          /* rank 0; 128 bytes */
int nbmf_comm_init(nbmf_ctx* ctx, const void* id128_host, int32_t rank, int32_t world);
int nbmf_comm_world(nbmf_ctx* ctx);
/* a communicator that outlives contexts (ncclCommInitRank costs ~1 s): create once per set of ranks, attach
 * to every context that shards over them, destroy after the last context is gone */
int nbmf_comm_create(const void* id128_host, int32_t rank, int32_t world, void** comm_out);
int nbmf_comm_attach(nbmf_ctx* ctx, void* comm, int32_t rank, int32_t world);
int nbmf_comm_destroy(void* comm);
/* engine actually selected for this context: NBMF_ENGINE_SIMT, NBMF_ENGINE_TENSOR or NBMF_ENGINE_FUSED */
int nbmf_engine(nbmf_ctx* ctx);
/* 1 when nbmf_fit / nbmf_fit_enqueue of this context run whole iterations inside the persistent small-fit kernel right
 * now (engine FUSED, no communicator attached, per-launch profiling off; env NBMF_NO_FUSED=1 at nbmf_create keeps AUTO
 * away from it): same arithmetic per entry as the SIMT pass kernels, one cooperative launch per nbmf_fit_enqueue call
 * instead of six launches per iteration (reference: the loop body of _solver.py:140-175) */
int nbmf_fit_is_fused(nbmf_ctx* ctx);

/* ---- the init stream of the reference for one row shard (host function, no GPU) ----
 * out[i] = lo + (hi - lo) * u_i for the `count` doubles that follow `skip` doubles in the stream of
 * numpy.random.RandomState(seed) -- the legacy MT19937 stream nbmf_mm_solver draws W_init (m x k) and then H_init (k x n)
 * from (_solver.py:102-103,126-129), bit for bit.  `skip` is reached by polynomial jump-ahead, not by drawing: a rank of
 * a row-sharded fit produces rows [r0, r1) of W_init (skip = r0 * k) and H_init (skip = m * k) without generating the
 * rows of the other ranks.  state_out (nullable, 625 x uint32): the generator state afterwards (key[624], pos), in the
 * layout numpy.random.set_state takes, so that a caller can leave the global stream where the reference leaves it. */
int nbmf_mt19937_uniform(uint32_t seed, uint64_t skip, uint64_t count, double lo, double hi, double* out_host,
                         uint32_t* state_out_host);

/* ---- measurement helpers ---- */
/* sustained FMA-pipe throughput (TFLOP/s) of packed FFMA2 (dtype F32) or DFMA (F64); synchronises */
int nbmf_fma_peak(int dtype, int32_t iters, void* scratch_dev, void* stream, double* tflops_host);
/* per-launch CUDA-event timing of the two pass kernels (H pass, W pass) on the context's stream:
 * enable clears the record; read synchronises and returns total milliseconds and launch counts */
int nbmf_profile_enable(nbmf_ctx* ctx, int enable);
int nbmf_profile_read(nbmf_ctx* ctx, double* h_ms_host, int32_t* h_count_host, double* w_ms_host, int32_t* w_count_host);
/* launch geometry chosen for this context: H pass = column blocks x row splits, W pass = row blocks x column splits */
int nbmf_plan_info(nbmf_ctx* ctx, int32_t* h_col_blocks, int32_t* h_row_splits, int32_t* w_row_blocks, int32_t* w_col_splits);
/* number of kernels this library launched since the last reset */
int64_t nbmf_launch_count(int reset);
/* tiling of the pass kernels chosen for (dtype, vkind, k): columns per H-pass CTA, rows per W-pass CTA, padded K */
int nbmf_variant_info(int dtype, int vkind, int k, int32_t* h_cols_per_cta, int32_t* w_rows_per_cta, int32_t* k_padded);

#ifdef __cplusplus
}
#endif
#endif /* NBMF_B200_H */

// C-ABI of libnbmf_b200.so (declared in include/nbmf_b200.h): variant dispatch, workspace
// planning, the device-resident fit loop and the NCCL row-shard reduction.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/nbmf_b200.h"
#include "internal.h"
#include "fused_args.h"

namespace nbmf {
bool lookup_f32_bits(int strict, int k, PassLaunch* out);
bool lookup_f32_dense(int strict, int k, PassLaunch* out);
bool lookup_f32_dense16(int strict, int k, PassLaunch* out);
bool lookup_f64_bits(int strict, int k, PassLaunch* out);
bool lookup_f64_dense(int strict, int k, PassLaunch* out);

bool lookup_pass(int dtype, int dense, int strict, int k, PassLaunch* out) {
  if (k < 1) return false;
  if (dtype == 0 && dense == 2) return lookup_f32_dense16(strict, k, out);
  if (dtype == 0) return dense ? lookup_f32_dense(strict, k, out) : lookup_f32_bits(strict, k, out);
  if (dtype == 1) return dense ? lookup_f64_dense(strict, k, out) : lookup_f64_bits(strict, k, out);
  return false;
}
}  // namespace nbmf

using namespace nbmf;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  return fail(NBMF_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CUDA_TRY(expr)                                         \
  do {                                                         \
    cudaError_t e__ = (expr);                                  \
    if (e__ != cudaSuccess) return cuda_fail(e__, #expr);      \
  } while (0)
#define CHECK_LAUNCH(n)                                        \
  do {                                                         \
    g_launches += (n);                                         \
    cudaError_t e__ = cudaPeekAtLastError();                   \
    if (e__ != cudaSuccess) return cuda_fail(e__, __func__);   \
  } while (0)

// ------------------------------------------------------------------------------------ NCCL (dlopen'd)
struct NcclId { char internal[128]; };
typedef int (*pfn_ncclGetUniqueId)(NcclId*);
typedef int (*pfn_ncclCommInitRank)(void**, int, NcclId, int);
typedef int (*pfn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*pfn_ncclCommDestroy)(void*);
typedef const char* (*pfn_ncclGetErrorString)(int);
typedef int (*pfn_ncclGroup)(void);
static struct {
  void* lib = nullptr;
  pfn_ncclGetUniqueId GetUniqueId = nullptr;
  pfn_ncclCommInitRank CommInitRank = nullptr;
  pfn_ncclAllReduce AllReduce = nullptr;
  pfn_ncclCommDestroy CommDestroy = nullptr;
  pfn_ncclGetErrorString GetErrorString = nullptr;
  pfn_ncclGroup GroupStart = nullptr, GroupEnd = nullptr;
} g_nccl;
constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

static int load_nccl() {
  if (g_nccl.AllReduce) return NBMF_OK;
  void* lib = nullptr;
  if (const char* p = getenv("NBMF_NCCL_LIB")) lib = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch already mapped
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(NBMF_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
  g_nccl.lib = lib;
  g_nccl.GetUniqueId = (pfn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (pfn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
  g_nccl.AllReduce = (pfn_ncclAllReduce)dlsym(lib, "ncclAllReduce");
  g_nccl.CommDestroy = (pfn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (pfn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
  g_nccl.GroupStart = (pfn_ncclGroup)dlsym(lib, "ncclGroupStart");
  g_nccl.GroupEnd = (pfn_ncclGroup)dlsym(lib, "ncclGroupEnd");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
    g_nccl.AllReduce = nullptr;
    return fail(NBMF_ERR_NCCL, "libnccl.so.2 lacks required symbols");
  }
  return NBMF_OK;
}
static int nccl_fail(int rc, const char* where) {
  return fail(NBMF_ERR_NCCL, std::string(where) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"));
}

// ------------------------------------------------------------------------------------ context
struct Plan {
  PassLaunch pl;
  int64_t ldh, wpr;
  int h_ncb, h_nsplit, w_nsplit, n_prior;
  int64_t h_rows_per_split, w_cols_per_split;
  size_t sz;   // sizeof(Real)
  bool tensor;           // tcgen05 engine (tc_passes.cuh) instead of the SIMT pass kernels
  int kb;                // tensor engine: K extent the MMAs cover (16 | 32 | 64); padded K (pl.kp) is 32 or 64
  int64_t mpad;          // tensor engine: rows padded to 128
  // workspace offsets
  size_t oW, oH, oHt, oCDpart, oCDsum, oLLpart, oLLsum, oPrior, oG, oQ, oRowcount, oHist, oState, oLoss, total;
  size_t oWf, oHf, oPc, oMc, oPM, oColcnt, oFlipcol, oFlipAny;   // tensor engine: formatted factor blocks, re-tiled bit planes, ones per column
  bool strict;
  // fused small-fit kernel (fused_small.cu): whole iterations in one persistent launch
  bool fused;
  int f_rows_per_warp, f_nsuper, f_nwords, f_h_in_smem;
  int f_wsplit;
  size_t f_smem, oFCD, oFLL, oBar, oFPrior2;
};

struct nbmf_ctx {
  nbmf_config cfg;
  Plan p;
  cudaStream_t st;
  unsigned char* ws;
  const uint32_t* P = nullptr;
  const uint32_t* M = nullptr;
  const void* Vm = nullptr;
  const void* Wm = nullptr;         // dense layout: values of a weighted (non-0/1) mask, or NULL
  bool rowcount_ready = false;
  int64_t ingest_rows = 0;          // rows handed over by nbmf_ingest_bits_rows so far
  // small problems: kGraphIters MM iterations captured once per fit into a CUDA graph (their ~7 launches per iteration
  // are what bounds many concurrent small fits); 0 = not built, 1 = ready, -1 = capture not possible on this stream
  int graph_state = 0;
  int fused_max_blocks = -1;        // co-resident CTAs of the fused small-fit kernel on this device (-1: not asked yet)
  unsigned long long* fused_trace = nullptr;   // development hook (env NBMF_FUSED_TRACE=1): phase time stamps, printed at destroy
  cudaGraphExec_t graph_exec = nullptr;
  long long graph_launches = 0;     // kernel launches per replay (for nbmf_launch_count)
  // batched small fits: this context leads batch_n contexts of identical configuration whose workspaces lie
  // batch_stride bytes apart; every launch of its fit loop covers all of them (gridDim.z), nbmf_batch_bind
  int batch_n = 1;
  int64_t batch_stride = 0;
  // loop
  int max_iter = 0;
  double tol = 0.0;
  int enqueued = 0;
  bool tail_enqueued = false;
  FitState* host_state = nullptr;   // pinned
  cudaEvent_t poll_ev = nullptr;
  bool poll_pending = false;
  // comm
  void* comm = nullptr;
  bool owns_comm = false;
  int world = 1, rank = 0;
  // optional per-launch timing of the two pass kernels (bench.py roofline)
  bool profile = false;
  std::vector<cudaEvent_t> prof_h, prof_w;      // (start, stop) pairs

  template <typename T> T* at(size_t off) const { return reinterpret_cast<T*>(ws + off); }
  void* W() const { return ws + p.oW; }
  void* H() const { return ws + p.oH; }
  void* Ht() const { return ws + p.oHt; }
  FitState* state() const { return at<FitState>(p.oState); }
};

static void prof_clear(std::vector<cudaEvent_t>& v);
static int format_w(nbmf_ctx* c, bool guarded);
static int format_h(nbmf_ctx* c, bool guarded);

static int choose_split(int64_t blocks, int64_t max_split, int occ = 1) {
  // fraction of the SM slots the grid keeps busy over its waves, times a mild preference for few splits
  // (every split costs a partial buffer and a term in the fixed-order reduction)
  if (max_split < 1) max_split = 1;
  const double sms = 148.0 * occ;
  int best = 1;
  double best_score = -1.0;
  for (int s = 1; s <= max_split; ++s) {
    const double ctas = (double)blocks * s, waves = ctas / sms;
    const double score = ctas / (ceil(waves) * sms) * (1.0 - 0.01 * s);
    if (score > best_score + 1e-9) { best = s; best_score = score; }
  }
  return best;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int make_plan(const nbmf_config& c, Plan* p) {
  if (c.m < 1 || c.n < 1) return fail(NBMF_ERR_ARG, "m and n must be positive");
  if (c.k < 1 || c.k > 128) return fail(NBMF_ERR_UNSUPPORTED, "n_components must be in 1..128");
  if (c.dtype != NBMF_F32 && c.dtype != NBMF_F64) return fail(NBMF_ERR_ARG, "dtype must be NBMF_F32 or NBMF_F64");
  if (c.vkind != NBMF_V_BITS && c.vkind != NBMF_V_DENSE && c.vkind != NBMF_V_DENSE_F16) return fail(NBMF_ERR_ARG, "bad vkind");
  if (c.vkind == NBMF_V_DENSE_F16 && c.dtype != NBMF_F32)
    return fail(NBMF_ERR_UNSUPPORTED, "the fp16 layout of probabilistic V needs float32 arithmetic");
  if (c.mask_semantics == NBMF_MASK_STRICT && !c.has_mask) { /* strict == reference when everything is observed */ }
  const int strict = (c.mask_semantics == NBMF_MASK_STRICT && c.has_mask) ? 1 : 0;
  if (!lookup_pass(c.dtype, c.vkind == NBMF_V_DENSE_F16 ? 2 : (c.vkind == NBMF_V_DENSE ? 1 : 0), strict, c.k, &p->pl))
    return fail(NBMF_ERR_UNSUPPORTED, "no kernel variant for this (dtype, vkind, k)");
  // engine: 0 = auto, 1 = SIMT (packed FFMA2), 2 = tensor (tcgen05, TF32 + bf16 split precision); NBMF_ENGINE overrides
  int engine = c.engine;
  if (const char* e = getenv("NBMF_ENGINE")) {
    if (!strcmp(e, "simt")) engine = NBMF_ENGINE_SIMT;
    else if (!strcmp(e, "tensor")) engine = NBMF_ENGINE_TENSOR;
    else if (!strcmp(e, "fused")) engine = NBMF_ENGINE_FUSED;
    else if (!strcmp(e, "auto")) engine = NBMF_ENGINE_AUTO;
  }
  // eps >= 1e-9: the tensor H pass takes one log per product of four x >= eps (no underflow)
  const bool eligible = c.dtype == NBMF_F32 && c.vkind == NBMF_V_BITS && c.k <= 64 && c.eps >= 1e-9;
  if (engine == NBMF_ENGINE_TENSOR && !eligible)
    return fail(NBMF_ERR_UNSUPPORTED, "tensor engine needs float32, bit-packed V, k <= 64 and eps >= 1e-9");
  p->strict = strict != 0;
  // auto rule: measured on the config-5 shape (1226 x 285, tools/small_fit_bench.py) the tensor engine wins from K = 16
  // for a single fit (12.6 vs 14.1 ms per 200 iterations; K = 6: 12.5 vs 10.8) and at every K for a batch of restarts
  int64_t min_m = 512, min_n = 128;
  if (const char* e = getenv("NBMF_TENSOR_MIN_M")) min_m = atoll(e);
  if (const char* e = getenv("NBMF_TENSOR_MIN_N")) min_n = atoll(e);
  // A single small fit is bound by launch latency on either engine: the fused small-fit kernel (fused_small.cu) takes
  // it when the work is small enough (measured crossover, tools/fused_sweep.py: entries x padded K of about 1.6e7 in fp64,
  // 1.2e7 in fp32 where the alternative is the tensor engine at 55-65 us per iteration).  NBMF_FUSED_MAX_WORK: experiment
  // knob; NBMF_NO_FUSED=1 keeps the launch-per-kernel paths.  fp32 takes one log per product of four x >= eps, fp64 per
  // eight: eps must keep the product a normal number.  An explicit engine = tensor wins (batches of fits: multifit.py).
  bool fused_fits = false;
  {
    int64_t max_work = c.dtype == NBMF_F32 ? (int64_t)12 << 20 : (int64_t)16 << 20;
    if (const char* e = getenv("NBMF_FUSED_MAX_WORK")) max_work = atoll(e);
    const bool no_fused = getenv("NBMF_NO_FUSED") != nullptr && strcmp(getenv("NBMF_NO_FUSED"), "0") != 0;
    const bool can_fuse = c.vkind == NBMF_V_BITS && c.k <= 32 && c.eps >= (c.dtype == NBMF_F32 ? 1e-9 : 1e-30);
    if (engine == NBMF_ENGINE_FUSED && !can_fuse)
      return fail(NBMF_ERR_UNSUPPORTED, "fused engine needs bit-packed V, k <= 32 and eps >= 1e-9 (float32) / 1e-30 (float64)");
    fused_fits = engine == NBMF_ENGINE_FUSED ||
                 (engine == NBMF_ENGINE_AUTO && !no_fused && can_fuse && c.m * c.n <= max_work / fused_kp(c.k));
  }
  p->tensor = eligible && (engine == NBMF_ENGINE_TENSOR ||
                           (engine == NBMF_ENGINE_AUTO && !fused_fits && c.m >= min_m && c.n >= min_n));
  p->kb = c.k <= 16 ? 16 : (c.k <= 32 ? 32 : 64);
  if (p->tensor) { p->pl.kp = c.k <= 32 ? 32 : 64; p->pl.h_bn = 128; p->pl.w_bmr = 128; }
  const int occ = 1;
  // batched small fits (nbmf_batch_bind): gridDim.z = bh fits share every launch, so the SMs are filled by the batch and
  // a fit should NOT be cut into many small row / column splits (each CTA pays the prologue that loads its factor
  // slice).  NBMF_BATCH_HINT is an experiment knob (tools/multifit_bench.py).
  int bh = std::max(1, (int)c.batch_hint);
  if (const char* e = getenv("NBMF_BATCH_HINT")) bh = std::max(1, atoi(e));
  p->sz = c.dtype == NBMF_F32 ? 4 : 8;
  p->wpr = nbmf_words_per_row(c.n);
  p->ldh = p->wpr * 32;
  const int kp = p->pl.kp;
  // H pass: column blocks x row splits
  p->h_ncb = (int)((c.n + p->pl.h_bn - 1) / p->pl.h_bn);
  // row splits of at least 128 rows, or 32 rows when the problem is too small to fill the GPU otherwise
  int64_t max_split = std::min<int64_t>(64, (c.m + 127) / 128);
  if ((int64_t)p->h_ncb * bh * max_split < 148) max_split = std::min<int64_t>(64, (c.m + 31) / 32);
  // tensor engine, small problems: a CTA pays ~4 us of set-up (TMEM allocation, barriers, resident operand, pipeline
  // fill) against ~0.4 us per 32-row block, so a row split is at least 512 rows -- also when that leaves SMs idle (a
  // batch of small fits fills them; the plan does not depend on the batch, so batched and single fits stay bit-identical)
  if (p->tensor && (int64_t)p->h_ncb * ((c.m + 127) / 128) < 148) max_split = std::max<int64_t>(1, std::min<int64_t>(64, c.m / 512));
  const size_t cd_one = (size_t)2 * kp * p->ldh * p->sz;
  while (max_split > 1 && cd_one * (size_t)max_split > ((size_t)2 << 30)) --max_split;
  p->h_nsplit = choose_split((int64_t)p->h_ncb * bh, max_split, occ);
  p->h_rows_per_split = ((c.m + p->h_nsplit - 1) / p->h_nsplit + 31) / 32 * 32;
  p->h_nsplit = (int)((c.m + p->h_rows_per_split - 1) / p->h_rows_per_split);
  // W pass: row blocks x column splits
  const int64_t nrb = (c.m + p->pl.w_bmr - 1) / p->pl.w_bmr;
  int64_t w_max_split = std::min<int64_t>(32, (c.n + 127) / 128);
  if (p->tensor && nrb * w_max_split < 148) w_max_split = std::max<int64_t>(1, std::min<int64_t>(32, c.n / 512));
  p->w_nsplit = choose_split(nrb * bh, w_max_split, occ);
  p->w_cols_per_split = ((c.n + p->w_nsplit - 1) / p->w_nsplit + 127) / 128 * 128;
  p->w_nsplit = (int)((c.n + p->w_cols_per_split - 1) / p->w_cols_per_split);
  p->n_prior = h_epilogue_blocks(c.n, kp);

  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  p->oW = take((size_t)c.m * kp * p->sz);
  p->oH = take((size_t)kp * p->ldh * p->sz);
  p->oHt = take((size_t)p->ldh * kp * p->sz);
  p->oCDpart = take(cd_one * p->h_nsplit);
  p->oCDsum = take(cd_one);
  p->oLLpart = take((size_t)p->h_nsplit * p->h_ncb * 8);
  p->oLLsum = take(64);
  p->oPrior = take((size_t)p->n_prior * 16);
  p->oG = take((size_t)p->w_nsplit * c.m * kp * p->sz);
  p->oQ = take((size_t)p->w_nsplit * c.m * p->sz);
  p->oRowcount = take((size_t)c.m * p->sz);
  p->oHist = take((size_t)(std::max(c.max_iter_cap, 1) + 2) * 8);
  p->oState = take(sizeof(FitState));
  p->oLoss = take(64);
  p->mpad = (c.m + 127) / 128 * 128;
  p->oWf = p->oHf = p->oPc = p->oMc = p->oPM = p->oColcnt = p->oFlipcol = p->oFlipAny = 0;
  if (p->tensor) {
    p->oWf = take((size_t)p->mpad * kp * 16);           // per row: hi + corr + transposed hi + corr = 4 x KT x 4 bytes
    p->oHf = take((size_t)p->ldh * kp * 16);
    p->oColcnt = take((size_t)p->ldh * 4);
    p->oFlipcol = take((size_t)p->ldh * 4);
    p->oFlipAny = take(64);
    p->oPc = take((size_t)p->ldh * p->mpad / 8);
    if (p->strict) p->oMc = take((size_t)p->ldh * p->mpad / 8);
    p->oPM = take((size_t)p->mpad * p->wpr * 8);
  }
  // fused small-fit kernel (decided above): its launch shape and workspace
  p->fused = false;
  p->oFCD = p->oFLL = p->oBar = p->oFPrior2 = 0;
  {
    if (fused_fits && !p->tensor) {
      // a CTA unit of the H phase = 8 warps x R rows x one 32-column word; about one unit per SM, 4 <= R <= 16
      const int nwords = (int)((c.n + 31) / 32);
      const int64_t target = std::max<int64_t>(1, 148 / nwords);
      const int64_t rows_per_unit = (c.m + target - 1) / target;
      const int R = (int)std::min<int64_t>(16, std::max<int64_t>(4, (rows_per_unit + 7) / 8));
      p->f_rows_per_warp = R;
      p->f_nsuper = (int)((c.m + 8 * R - 1) / (8 * R));
      p->f_nwords = nwords;
      p->f_h_in_smem = (size_t)c.k * nwords * 32 * p->sz <= (size_t)160 * 1024 ? 1 : 0;
      p->f_smem = fused_smem_bytes(c.dtype, c.k, R, nwords, p->f_h_in_smem);
      p->fused = p->f_smem <= (size_t)200 * 1024;
      if (p->fused) {
        p->oFCD = take((size_t)p->f_nsuper * 2 * kp * nwords * 32 * p->sz);
        p->oFLL = take((size_t)nwords * p->f_nsuper * 8);
        p->oBar = take(64);
        p->oFPrior2 = take((size_t)p->n_prior * 16);
        // W phase: warps that share a row, so that a problem with few rows still uses every warp of the (nominal) grid
        // (more warps per row also when the rows exceed the grid's warps -- 1226 rows: 3 trips of 5 words instead of 2 of
        // 9 -- measured slower: every trip pays two CTA barriers)
        int ws = 1;
        while (ws < 8 && (int64_t)ws * 2 * c.m <= 148 * 8 && ws * 2 <= nwords) ws *= 2;
        p->f_wsplit = ws;
      }
    }
  }
  p->total = o;
  return NBMF_OK;
}

// Pinned FitState snapshots are pooled: cudaMallocHost / cudaFreeHost cost ~0.1-1 ms and synchronise the device,
// which matters when many small fits are created and destroyed (restarts, grids) or run concurrently.
#include <mutex>
static std::mutex g_pin_mu;
static std::vector<FitState*> g_pin_free;
static FitState* pinned_state_get() {
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (!g_pin_free.empty()) { FitState* s = g_pin_free.back(); g_pin_free.pop_back(); return s; }
  }
  FitState* block = nullptr;
  constexpr int kSlots = 64;
  if (cudaMallocHost((void**)&block, sizeof(FitState) * kSlots) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  for (int i = 1; i < kSlots; ++i) g_pin_free.push_back(block + i);
  return block;
}
static void pinned_state_put(FitState* s) {
  if (!s) return;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_pin_free.push_back(s);
}

// ------------------------------------------------------------------------------------ small API
extern "C" int nbmf_version(void) { return 100; }
extern "C" const char* nbmf_last_error(void) { return g_err.c_str(); }
extern "C" int64_t nbmf_words_per_row(int64_t n) { return (n + 1023) / 1024 * 32; }
extern "C" int64_t nbmf_padded_cols(int64_t n) { return nbmf_words_per_row(n) * 32; }
extern "C" int64_t nbmf_launch_count(int reset) {
  const long long v = g_launches.load();
  if (reset) g_launches = 0;
  return v;
}
extern "C" int nbmf_variant_info(int dtype, int vkind, int k, int32_t* h_cols, int32_t* w_rows, int32_t* kp) {
  PassLaunch pl;
  if (!lookup_pass(dtype, vkind == NBMF_V_DENSE_F16 ? 2 : (vkind == NBMF_V_DENSE ? 1 : 0), 0, k, &pl)) return fail(NBMF_ERR_UNSUPPORTED, "no variant");
  if (h_cols) *h_cols = pl.h_bn;
  if (w_rows) *w_rows = pl.w_bmr;
  if (kp) *kp = pl.kp;
  return NBMF_OK;
}

// ------------------------------------------------------------------------------------ data layer
extern "C" int nbmf_pack_bits(const void* x, int xdt, int64_t ldx, const void* mask, int mdt, int64_t ldm, int64_t m,
                              int64_t n, uint32_t* P, uint32_t* M, void* stream) {
  if (!x || !P || m < 1 || n < 1) return fail(NBMF_ERR_ARG, "nbmf_pack_bits: bad arguments");
  launch_pack_bits(xdt, x, ldx, mask, mdt, ldm, m, n, nbmf_words_per_row(n), P, M, nullptr, (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}
extern "C" int nbmf_pack_bits_checked(const void* x, int xdt, int64_t ldx, const void* mask, int mdt, int64_t ldm, int64_t m,
                                      int64_t n, uint32_t* P, uint32_t* M, int32_t* flags_dev, void* stream) {
  if (!x || !P || !flags_dev || m < 1 || n < 1) return fail(NBMF_ERR_ARG, "nbmf_pack_bits_checked: bad arguments");
  launch_pack_bits(xdt, x, ldx, mask, mdt, ldm, m, n, nbmf_words_per_row(n), P, M, flags_dev, (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}
extern "C" int nbmf_pack_dense(const void* x, int xdt, int64_t ldx, const void* mask, int mdt, int64_t ldm, int64_t m,
                               int64_t n, int out_dtype, void* vm, void* stream) {
  if (!x || !vm || m < 1 || n < 1) return fail(NBMF_ERR_ARG, "nbmf_pack_dense: bad arguments");
  launch_pack_dense(xdt, x, ldx, mask, mdt, ldm, m, n, out_dtype, nbmf_padded_cols(n), vm, (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}
extern "C" int nbmf_pack_csr(const int64_t* indptr, const int32_t* indices, const void* data, int data_dtype, int64_t m,
                             int64_t n, uint32_t* P, int32_t* flags_dev, int32_t* flags_host, void* stream) {
  if (!indptr || !P || !flags_dev || !flags_host || m < 1 || n < 1)     // indices may be NULL when nnz == 0
    return fail(NBMF_ERR_ARG, "nbmf_pack_csr: bad arguments");
  if (data && data_dtype != NBMF_F32 && data_dtype != NBMF_F64) return fail(NBMF_ERR_ARG, "nbmf_pack_csr: data must be f32 or f64");
  launch_pack_csr(indptr, indices, data, data_dtype, m, n, nbmf_words_per_row(n), P, flags_dev, (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  CUDA_TRY(cudaMemcpyAsync(flags_host, flags_dev, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return NBMF_OK;
}
extern "C" int nbmf_reconstruct(int dtype, const void* w, const void* h, int64_t m, int64_t n, int32_t k, void* out, void* stream) {
  if (!w || !h || !out || m < 1 || n < 1 || k < 1 || k > 128) return fail(NBMF_ERR_ARG, "nbmf_reconstruct: bad arguments (k in 1..128)");
  if (dtype != NBMF_F32 && dtype != NBMF_F64) return fail(NBMF_ERR_ARG, "dtype must be NBMF_F32 or NBMF_F64");
  launch_reconstruct(dtype, w, h, m, n, k, out, (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}
extern "C" int nbmf_transpose_bits(const uint32_t* src, int64_t m, int64_t n, uint32_t* dst, void* stream) {
  if (!src || !dst || m < 1 || n < 1) return fail(NBMF_ERR_ARG, "nbmf_transpose_bits: bad arguments");
  launch_transpose_bits(src, m, n, nbmf_words_per_row(n), dst, nbmf_words_per_row(m), (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}
extern "C" int nbmf_popcount_bits(const uint32_t* bits, int64_t m, int64_t n, uint64_t* scratch, uint64_t* count_host,
                                  void* stream) {
  if (!bits || !scratch || !count_host) return fail(NBMF_ERR_ARG, "nbmf_popcount_bits: bad arguments");
  launch_popcount(bits, m, nbmf_words_per_row(n), (unsigned long long*)scratch, (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  CUDA_TRY(cudaMemcpyAsync(count_host, scratch, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return NBMF_OK;
}
extern "C" int nbmf_synth_bits(uint64_t seed, int64_t row0, int64_t m, int64_t n, const float* hstar, int32_t kstar,
                               float obs_frac, uint32_t* P, uint32_t* M, void* stream) {
  if (!hstar || !P || kstar < 1 || kstar > 32) return fail(NBMF_ERR_ARG, "nbmf_synth_bits: bad arguments (kstar in 1..32)");
  launch_synth_bits(seed, row0, m, n, nbmf_words_per_row(n), nullptr, hstar, kstar, obs_frac, P, M, (cudaStream_t)stream);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}

// ------------------------------------------------------------------------------------ context
extern "C" int64_t nbmf_workspace_bytes(const nbmf_config* cfg) {
  Plan p;
  if (!cfg) { fail(NBMF_ERR_ARG, "null config"); return -1; }
  if (make_plan(*cfg, &p) != NBMF_OK) return -1;
  return (int64_t)p.total;
}

__global__ void reset_state_kernel(FitState* s, double alpha, double beta) {
  s->done = 0; s->it = 0; s->n_hist = 0; s->converged = 0;
  s->prev_loss = INFINITY; s->prior_a = 0.0; s->prior_b = 0.0;
  s->alpha = alpha; s->beta = beta;
}

extern "C" int nbmf_create(const nbmf_config* cfg, void* ws, int64_t ws_bytes, void* stream, nbmf_ctx** out) {
  if (!cfg || !ws || !out) return fail(NBMF_ERR_ARG, "nbmf_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(NBMF_ERR_CUDA, "no CUDA device: libnbmf_b200 has no CPU fallback");
  nbmf_ctx* c = new nbmf_ctx();
  c->cfg = *cfg;
  int rc = make_plan(*cfg, &c->p);
  if (rc != NBMF_OK) { delete c; return rc; }
  if ((size_t)ws_bytes < c->p.total) { delete c; return fail(NBMF_ERR_ARG, "workspace too small"); }
  if ((uintptr_t)ws % 256) { delete c; return fail(NBMF_ERR_ARG, "workspace must be 256-byte aligned"); }
  c->ws = (unsigned char*)ws;
  c->st = (cudaStream_t)stream;
  c->host_state = pinned_state_get();
  if (!c->host_state) { delete c; return cuda_fail(cudaGetLastError(), "cudaMallocHost"); }
  cudaError_t e = cudaEventCreateWithFlags(&c->poll_ev, cudaEventDisableTiming);
  if (e != cudaSuccess) { pinned_state_put(c->host_state); delete c; return cuda_fail(e, "cudaEventCreate"); }
  reset_state_kernel<<<1, 1, 0, c->st>>>(c->state(), c->cfg.alpha, c->cfg.beta);
  g_launches += 1;
  if (c->p.tensor) {                                   // "no column flips" until flip_cols_kernel says otherwise
    cudaMemsetAsync(c->at<uint32_t>(c->p.oFlipcol), 0, (size_t)c->p.ldh * 4, c->st);
    cudaMemsetAsync(c->at<int>(c->p.oFlipAny), 0, 64, c->st);
  }
  *out = c;
  return NBMF_OK;
}

static void graph_drop(nbmf_ctx* c) {
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  c->graph_exec = nullptr;
  c->graph_state = 0;
}

extern "C" int nbmf_destroy(nbmf_ctx* c) {
  if (!c) return NBMF_OK;
  cudaStreamSynchronize(c->st);     // the caller frees the workspace next (not a device-wide sync: other streams may be capturing)
  if (c->fused_trace) {             // mean nanoseconds per phase of CTA 0 over the traced iterations
    std::vector<unsigned long long> t(4096 * 16);
    if (cudaMemcpy(t.data(), c->fused_trace, t.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
      double sum[8] = {0}, sub[6] = {0}; int cnt = 0;
      for (int it = 1; it < 4096; ++it) {
        const unsigned long long* r = &t[(size_t)it * 16];
        if (!r[0] || !r[7]) continue;
        for (int s = 1; s < 8; ++s) sum[s] += (double)(r[s] - r[s - 1]);
        sub[0] += (double)(r[8] - r[0]); sub[1] += (double)(r[9] - r[8]); sub[2] += (double)(r[10] - r[9]);
        sub[3] += (double)(r[11] - r[5]); sub[4] += (double)(r[12] - r[11]); sub[5] += (double)(r[13] - r[12]);
        ++cnt;
      }
      if (cnt) {
        fprintf(stderr, "[nbmf fused trace] %d iterations, ns: H phase %.0f | barrier %.0f | finalize %.0f | H epilogue %.0f | barrier %.0f | W phase %.0f | barrier %.0f\n",
                cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt, sum[4] / cnt, sum[5] / cnt, sum[6] / cnt, sum[7] / cnt);
        fprintf(stderr, "[nbmf fused trace]   H phase: stage %.0f, rows %.0f, reduce+store %.0f; W phase: stage %.0f, words %.0f, warp sums %.0f\n",
                sub[0] / cnt, sub[1] / cnt, sub[2] / cnt, sub[3] / cnt, sub[4] / cnt, sub[5] / cnt);
      }
    }
    cudaFree(c->fused_trace);
  }
  graph_drop(c);
  if (c->comm && c->owns_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  prof_clear(c->prof_h);
  prof_clear(c->prof_w);
  if (c->poll_ev) cudaEventDestroy(c->poll_ev);
  pinned_state_put(c->host_state);
  delete c;
  return NBMF_OK;
}

// Tensor engine: the pass kernels read the re-tiled copies only; the one remaining reader of the caller's row-major
// mask plane is the per-row observed count of the Duchi projection, taken here, so that the caller may free or reuse
// its planes right after handing them over (nbmf_planes_in_use): 25 GB of 62.5 GB at config 4.
static int eager_rowcount(nbmf_ctx* c) {
  if (c->cfg.projection == NBMF_PROJ_DUCHI && c->M && !c->rowcount_ready) {
    launch_rowcount(c->cfg.dtype, c->M, c->cfg.m, c->cfg.n, c->p.wpr, c->ws + c->p.oRowcount, c->st);
    CHECK_LAUNCH(1);
    c->rowcount_ready = true;
  }
  return NBMF_OK;
}
// 1 while the context still reads the planes given to nbmf_set_data_bits / nbmf_ingest_bits_begin (SIMT engine: always;
// tensor engine: only until the work enqueued by nbmf_set_data_bits / nbmf_ingest_bits_end has run -- stream order makes
// a free or overwrite enqueued on the same stream afterwards safe).
extern "C" int nbmf_planes_in_use(nbmf_ctx* c) {
  if (!c) return 0;
  if (c->cfg.vkind != NBMF_V_BITS) return 1;
  if (!c->p.tensor) return 1;
  return (c->ingest_rows != 0 && c->ingest_rows != c->cfg.m) ? 1 : 0;     // a streamed ingestion is still in progress
}

extern "C" int nbmf_set_data_bits(nbmf_ctx* c, const uint32_t* P, const uint32_t* M) {
  if (!c || !P) return fail(NBMF_ERR_ARG, "nbmf_set_data_bits: null argument");
  if (c->cfg.vkind != NBMF_V_BITS) return fail(NBMF_ERR_ARG, "context was created for dense V");
  if (c->cfg.has_mask && !M) return fail(NBMF_ERR_ARG, "has_mask is set but no mask plane given");
  c->P = P;
  c->M = c->cfg.has_mask ? M : nullptr;
  c->rowcount_ready = false;
  graph_drop(c);
  if (c->p.tensor) {   // the tensor kernels read planes re-tiled per TMEM lane (format_factors.cu)
    launch_tile_planes(P, c->M, c->cfg.m, c->cfg.n, c->p.wpr, c->p.mpad, 0, c->cfg.m, c->at<uint32_t>(c->p.oPc),
                       c->p.strict ? c->at<uint32_t>(c->p.oMc) : nullptr, c->ws + c->p.oPM, c->st);
    CUDA_TRY(cudaMemsetAsync(c->at<uint32_t>(c->p.oColcnt), 0, (size_t)c->p.ldh * 4, c->st));
    launch_colcount(c->at<uint32_t>(c->p.oPc), c->p.mpad / 32, 0, c->p.mpad / 32, c->p.ldh, c->at<uint32_t>(c->p.oColcnt), c->st);
    CHECK_LAUNCH(c->p.strict ? 4 : 3);
    int rc = eager_rowcount(c);
    if (rc) return rc;
  }
  return NBMF_OK;
}

// Streamed ingestion: the caller copies the host planes to the device chunk by chunk on its own copy stream, orders
// the context's stream after each chunk (event) and calls nbmf_ingest_bits_rows for it, so that P &= M, the mask
// count and the re-tiling for the tensor engine run while later chunks are still crossing PCIe.
static unsigned long long* ingest_counter(nbmf_ctx* c) { return c->at<unsigned long long>(c->p.oLoss) + 4; }
extern "C" int nbmf_ingest_bits_begin(nbmf_ctx* c, uint32_t* P, const uint32_t* M) {
  if (!c || !P) return fail(NBMF_ERR_ARG, "nbmf_ingest_bits_begin: null argument");
  if (c->cfg.vkind != NBMF_V_BITS) return fail(NBMF_ERR_ARG, "context was created for dense V");
  if (c->cfg.has_mask && !M) return fail(NBMF_ERR_ARG, "has_mask is set but no mask plane given");
  c->P = P;
  c->M = c->cfg.has_mask ? M : nullptr;
  c->rowcount_ready = false;
  graph_drop(c);
  c->ingest_rows = 0;
  CUDA_TRY(cudaMemsetAsync(ingest_counter(c), 0, sizeof(unsigned long long), c->st));
  if (c->p.tensor) CUDA_TRY(cudaMemsetAsync(c->at<uint32_t>(c->p.oColcnt), 0, (size_t)c->p.ldh * 4, c->st));
  return NBMF_OK;
}
extern "C" int nbmf_ingest_bits_rows(nbmf_ctx* c, int64_t row0, int64_t row1) {
  if (!c || !c->P) return fail(NBMF_ERR_ARG, "nbmf_ingest_bits_begin was not called");
  if (row0 != c->ingest_rows || row1 <= row0 || row1 > c->cfg.m || (row0 % 128) != 0)
    return fail(NBMF_ERR_ARG, "nbmf_ingest_bits_rows: chunks must be consecutive, start on multiples of 128 rows and end at <= m");
  uint32_t* P = const_cast<uint32_t*>(c->P);
  int launches = 0;
  if (c->M) {
    launch_and_count(P + row0 * c->p.wpr, c->M + row0 * c->p.wpr, (row1 - row0) * c->p.wpr, ingest_counter(c), c->st);
    ++launches;
  }
  if (c->p.tensor) {
    launch_tile_planes(P, c->M, c->cfg.m, c->cfg.n, c->p.wpr, c->p.mpad, row0, row1, c->at<uint32_t>(c->p.oPc),
                       c->p.strict ? c->at<uint32_t>(c->p.oMc) : nullptr, c->ws + c->p.oPM, c->st);
    const int64_t end = row1 >= c->cfg.m ? c->p.mpad : row1;         // the re-tiled words of this chunk exist now
    launch_colcount(c->at<uint32_t>(c->p.oPc), c->p.mpad / 32, row0 / 32, (end + 31) / 32, c->p.ldh,
                    c->at<uint32_t>(c->p.oColcnt), c->st);
    launches += c->p.strict ? 4 : 3;
  }
  CHECK_LAUNCH(launches);
  c->ingest_rows = row1;
  return NBMF_OK;
}
extern "C" int nbmf_ingest_bits_end(nbmf_ctx* c, double* mask_count_host) {
  if (!c || !c->P) return fail(NBMF_ERR_ARG, "nbmf_ingest_bits_begin was not called");
  if (c->ingest_rows != c->cfg.m) return fail(NBMF_ERR_ARG, "nbmf_ingest_bits_end: not all rows were ingested");
  if (c->p.tensor) {
    int rc = eager_rowcount(c);
    if (rc) return rc;
  }
  unsigned long long h = 0;
  CUDA_TRY(cudaMemcpyAsync(&h, ingest_counter(c), 8, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  if (mask_count_host) *mask_count_host = c->M ? (double)h : (double)c->cfg.m * (double)c->cfg.n;
  return NBMF_OK;
}
extern "C" int nbmf_set_n_obs(nbmf_ctx* c, double n_obs) {
  if (!c || !(n_obs > 0)) return fail(NBMF_ERR_ARG, "nbmf_set_n_obs: bad arguments");
  c->cfg.n_obs = n_obs;
  graph_drop(c);
  return NBMF_OK;
}
extern "C" int nbmf_set_data_dense(nbmf_ctx* c, const void* Vm, const uint32_t* M) {
  if (!c || !Vm) return fail(NBMF_ERR_ARG, "nbmf_set_data_dense: null argument");
  if (c->cfg.vkind != NBMF_V_DENSE && c->cfg.vkind != NBMF_V_DENSE_F16) return fail(NBMF_ERR_ARG, "context was created for bit-packed V");
  if (c->cfg.has_mask && !M) return fail(NBMF_ERR_ARG, "has_mask is set but no mask plane given");
  c->Vm = Vm;
  c->Wm = nullptr;
  c->M = c->cfg.has_mask ? M : nullptr;
  c->rowcount_ready = false;
  graph_drop(c);
  return NBMF_OK;
}

// Weighted observation mask (values other than 0 / 1): the reference multiplies by the mask values (Y * mask,
// (1 - Y).T * mask.T: _solver.py:30-32).  V * mask is what nbmf_set_data_dense takes anyway; the W half-step additionally
// needs the mask values themselves ((1 - V) * mask = mask - V * mask), in the same dense layout and dtype.  The bit plane
// given to nbmf_set_data_dense is then (mask != 0).  Reference mask semantics only.
extern "C" int nbmf_set_mask_weights(nbmf_ctx* c, const void* wm) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  if (c->cfg.vkind != NBMF_V_DENSE) return fail(NBMF_ERR_UNSUPPORTED, "weighted masks need the dense V layout in the compute dtype");
  if (wm && c->p.strict) return fail(NBMF_ERR_UNSUPPORTED, "weighted masks are defined for the reference mask semantics only");
  c->Wm = wm;
  graph_drop(c);
  return NBMF_OK;
}

static int require_data(nbmf_ctx* c) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  if (c->cfg.vkind == NBMF_V_BITS ? !c->P : !c->Vm) return fail(NBMF_ERR_ARG, "no data planes set");
  return NBMF_OK;
}

extern "C" int nbmf_set_factors(nbmf_ctx* c, const void* w, const void* h, int normalize_w) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  launch_init_factors(c->cfg.dtype, w, h, c->cfg.m, c->cfg.n, c->cfg.k, c->p.pl.kp, c->p.ldh, c->W(), c->H(), c->Ht(),
                      normalize_w, c->st);
  reset_state_kernel<<<1, 1, 0, c->st>>>(c->state(), c->cfg.alpha, c->cfg.beta);
  CHECK_LAUNCH((w ? 1 : 0) + (h ? 1 : 0) + 1);
  c->enqueued = 0;
  c->tail_enqueued = false;
  int rc;
  if (w && (rc = format_w(c, false))) return rc;
  if (h && (rc = format_h(c, false))) return rc;
  return NBMF_OK;
}
extern "C" int nbmf_get_factors(nbmf_ctx* c, void* w, void* h) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  launch_export_factors(c->cfg.dtype, c->W(), c->H(), c->cfg.m, c->cfg.n, c->cfg.k, c->p.pl.kp, c->p.ldh, w, h, c->st);
  CHECK_LAUNCH((w ? 1 : 0) + (h ? 1 : 0));
  return NBMF_OK;
}

extern "C" int nbmf_simplex_deviation(nbmf_ctx* c, double* dev_host) {
  if (!c || !dev_host) return fail(NBMF_ERR_ARG, "nbmf_simplex_deviation: null argument");
  unsigned long long* scratch = c->at<unsigned long long>(c->p.oLoss);     // 64 bytes of scratch
  launch_simplex_deviation(c->cfg.dtype, c->W(), c->cfg.m, c->cfg.k, c->p.pl.kp, scratch, c->st);
  CHECK_LAUNCH(1);
  unsigned long long h[2];
  CUDA_TRY(cudaMemcpyAsync(h, scratch, 16, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  double d;
  memcpy(&d, &h[0], 8);
  *dev_host = h[1] ? NAN : d;
  return NBMF_OK;
}
extern "C" int nbmf_get_factors_f64(nbmf_ctx* c, double* w, double* h, int normalize_w) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  launch_export_f64(c->cfg.dtype, c->W(), c->H(), c->cfg.m, c->cfg.n, c->cfg.k, c->p.pl.kp, c->p.ldh, normalize_w, w, h, c->st);
  CHECK_LAUNCH((w ? 1 : 0) + (h ? 1 : 0));
  return NBMF_OK;
}

// ------------------------------------------------------------------------------------ steps
static int format_w(nbmf_ctx* c, bool guarded) {
  if (!c->p.tensor) return NBMF_OK;
  launch_format_w(c->W(), c->cfg.m, c->p.mpad, c->p.pl.kp, c->ws + c->p.oWf, guarded ? c->state() : nullptr, c->st,
                  guarded ? c->batch_n : 1, guarded ? c->batch_stride : 0);     // unguarded = per-context set-up calls
  CHECK_LAUNCH(1);
  return NBMF_OK;
}
static int format_h(nbmf_ctx* c, bool guarded) {
  if (!c->p.tensor) return NBMF_OK;
  // K <= 32 kernels: MMA1 of the W pass yields Theta + eps (see format_h_kernel); the K <= 64 kernels add eps themselves
  const float theta_bias = c->p.pl.kp == 32 ? (float)c->cfg.eps : 0.0f;
  launch_format_h(c->H(), c->p.ldh, c->p.pl.kp, theta_bias, c->ws + c->p.oHf, guarded ? c->state() : nullptr, c->st,
                  guarded ? c->batch_n : 1, guarded ? c->batch_stride : 0);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}

static int allreduce(nbmf_ctx* c, bool with_cd) {
  if (c->world <= 1) return NBMF_OK;
  const size_t count = (size_t)2 * c->p.pl.kp * c->p.ldh;
  int rc = g_nccl.GroupStart();
  if (rc) return nccl_fail(rc, "ncclGroupStart");
  if (with_cd) {
    void* cd = c->ws + c->p.oCDsum;
    rc = g_nccl.AllReduce(cd, cd, count, c->cfg.dtype == NBMF_F32 ? kNcclFloat32 : kNcclFloat64, kNcclSum, c->comm, c->st);
    if (rc) return nccl_fail(rc, "ncclAllReduce(C|D)");
  }
  void* ll = c->ws + c->p.oLLsum;
  rc = g_nccl.AllReduce(ll, ll, 1, kNcclFloat64, kNcclSum, c->comm, c->st);
  if (rc) return nccl_fail(rc, "ncclAllReduce(LL)");
  rc = g_nccl.GroupEnd();
  if (rc) return nccl_fail(rc, "ncclGroupEnd");
  return NBMF_OK;
}

static void prof_mark(nbmf_ctx* c, std::vector<cudaEvent_t>& v) {
  if (!c->profile) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, c->st);
  v.push_back(e);
}

static FinalizeArgs finalize_args(nbmf_ctx* c) {
  const Plan& p = c->p;
  return FinalizeArgs{c->state(), c->at<double>(p.oPrior), p.n_prior, c->cfg.alpha, c->cfg.beta, c->cfg.n_obs, c->tol,
                      c->max_iter, c->at<double>(p.oHist)};
}

// with_finalize: the loss / stop rule of the fit loop follows this pass.  On one GPU it rides in the reduction kernel;
// with row shards the log-likelihood is all-reduced first and finalize is its own launch.
static int enqueue_h_pass(nbmf_ctx* c, int compute_cd, bool with_finalize = false) {
  const Plan& p = c->p;
  HPassArgs a;
  a.W = c->W(); a.H = c->H(); a.P = c->P; a.M = c->M; a.Vm = c->Vm; a.ldv = p.ldh;
  a.m = c->cfg.m; a.n = c->cfg.n; a.ldh = p.ldh; a.wpr = p.wpr;
  a.rows_per_split = p.h_rows_per_split;
  a.CD = c->ws + p.oCDpart; a.LL = c->at<double>(p.oLLpart);
  a.eps = c->cfg.eps; a.done = &c->state()->done; a.compute_cd = compute_cd;
  a.batch_n = c->batch_n; a.batch_stride = c->batch_stride;
  prof_mark(c, c->prof_h);
  if (p.tensor)
    launch_h_pass_tensor(a, c->ws + p.oWf, c->at<uint32_t>(p.oPc), p.strict ? c->at<uint32_t>(p.oMc) : nullptr, p.mpad / 32,
                         p.kb, c->cfg.k, getenv("NBMF_TC_NOFLIP") ? nullptr : c->at<uint32_t>(p.oColcnt),
                         c->at<uint32_t>(p.oFlipcol), c->at<int>(p.oFlipAny), p.h_nsplit, c->st);
  else
    p.pl.h_launch(a, p.h_nsplit, c->st);
  prof_mark(c, c->prof_h);
  const int64_t count = compute_cd ? (int64_t)2 * p.pl.kp * p.ldh : 0;
  const bool fused = with_finalize && c->world == 1;
  FinalizeArgs fin = finalize_args(c);
  if (!fused) fin.state = nullptr;
  launch_h_reduce(c->cfg.dtype, c->ws + p.oCDpart, p.h_nsplit, count, c->ws + p.oCDsum, c->at<double>(p.oLLpart),
                  (int64_t)p.h_nsplit * p.h_ncb, c->at<double>(p.oLLsum), c->state(), fin, c->st, c->batch_n, c->batch_stride);
  CHECK_LAUNCH(2);
  int rc = allreduce(c, compute_cd != 0);
  if (rc || !with_finalize || fused) return rc;
  launch_finalize(finalize_args(c), c->at<double>(p.oLLsum), c->st);
  CHECK_LAUNCH(1);
  return NBMF_OK;
}

static int enqueue_h_epilogue(nbmf_ctx* c) {
  const Plan& p = c->p;
  launch_h_epilogue(c->cfg.dtype, c->ws + p.oCDsum, c->cfg.n, c->cfg.k, p.pl.kp, p.ldh, c->cfg.alpha, c->cfg.beta,
                    c->cfg.eps, c->H(), c->Ht(), c->at<double>(p.oPrior), c->state(), c->st, c->batch_n, c->batch_stride);
  CHECK_LAUNCH(1);
  return format_h(c, true);
}

static int enqueue_w_step(nbmf_ctx* c) {
  const Plan& p = c->p;
  if (c->cfg.projection == NBMF_PROJ_DUCHI && c->M && !c->rowcount_ready) {
    launch_rowcount(c->cfg.dtype, c->M, c->cfg.m, c->cfg.n, p.wpr, c->ws + p.oRowcount, c->st);
    CHECK_LAUNCH(1);
    c->rowcount_ready = true;
  }
  WPassArgs a;
  a.W = c->W(); a.Ht = c->Ht(); a.P = c->P; a.M = c->M; a.Vm = c->Vm; a.Wm = c->Wm; a.ldv = p.ldh;
  a.m = c->cfg.m; a.n = c->cfg.n; a.ldh = p.ldh; a.wpr = p.wpr;
  a.cols_per_split = p.w_cols_per_split;
  a.G = c->ws + p.oG; a.Q = c->ws + p.oQ; a.eps = c->cfg.eps; a.done = &c->state()->done;
  a.batch_n = c->batch_n; a.batch_stride = c->batch_stride;
  prof_mark(c, c->prof_w);
  if (p.tensor)
    launch_w_pass_tensor(a, c->ws + p.oHf, c->ws + p.oPM, p.kb, p.w_nsplit, c->st);
  else
    p.pl.w_launch(a, p.w_nsplit, c->st);
  prof_mark(c, c->prof_w);
  const void* rowcount = (c->cfg.projection == NBMF_PROJ_DUCHI && c->M) ? (const void*)(c->ws + p.oRowcount) : nullptr;
  launch_w_epilogue(c->cfg.dtype, c->ws + p.oG, c->ws + p.oQ, p.w_nsplit, c->cfg.m, c->cfg.n, c->cfg.k, p.pl.kp,
                    c->cfg.projection, rowcount, c->W(), c->state(), c->st, c->batch_n, c->batch_stride);
  CHECK_LAUNCH(2);
  return format_w(c, true);
}

extern "C" int nbmf_h_half_step(nbmf_ctx* c) {
  int rc = require_data(c);
  if (rc) return rc;
  if ((rc = enqueue_h_pass(c, 1))) return rc;
  return enqueue_h_epilogue(c);
}
extern "C" int nbmf_w_half_step(nbmf_ctx* c) {
  int rc = require_data(c);
  if (rc) return rc;
  return enqueue_w_step(c);
}

__global__ void objective_kernel(const double* __restrict__ LLsum, const double* __restrict__ prior_part, int n_part,
                                 double alpha, double beta, double n_obs, double* __restrict__ out) {
  double pa = 0.0, pb = 0.0;
  for (int i = threadIdx.x; i < n_part; i += 32) { pa += prior_part[2 * i]; pb += prior_part[2 * i + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pa += __shfl_xor_sync(0xffffffffu, pa, o);
    pb += __shfl_xor_sync(0xffffffffu, pb, o);
  }
  if (threadIdx.x == 0) out[0] = -(LLsum[0] + (alpha - 1.0) * pa + (beta - 1.0) * pb) / n_obs;
}

extern "C" int nbmf_objective(nbmf_ctx* c, double* loss_host) {
  int rc = require_data(c);
  if (rc) return rc;
  if (!loss_host) return fail(NBMF_ERR_ARG, "null output");
  const Plan& p = c->p;
  if ((rc = enqueue_h_pass(c, 0))) return rc;
  launch_prior_sums(c->cfg.dtype, c->H(), c->cfg.n, c->cfg.k, p.pl.kp, p.ldh, c->cfg.eps, c->at<double>(p.oPrior), c->st);
  objective_kernel<<<1, 32, 0, c->st>>>(c->at<double>(p.oLLsum), c->at<double>(p.oPrior), p.n_prior, c->cfg.alpha,
                                        c->cfg.beta, c->cfg.n_obs, c->at<double>(p.oLoss));
  CHECK_LAUNCH(2);
  CUDA_TRY(cudaMemcpyAsync(loss_host, c->at<double>(p.oLoss), 8, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  return NBMF_OK;
}

// Per-CTA log-likelihood partials of the most recent H pass (row split s, column block b at [s * col_blocks + b]):
// lets a caller check the fused NLL of one column block against a host computation when the whole matrix is far too
// large for one (bench.py parity_check at config 4).
extern "C" int nbmf_loglik_partials(nbmf_ctx* c, double* out_host, int64_t capacity, int32_t* col_blocks, int32_t* row_splits) {
  if (!c || !out_host) return fail(NBMF_ERR_ARG, "nbmf_loglik_partials: null argument");
  if (c->batch_n > 1) return fail(NBMF_ERR_ARG, "nbmf_loglik_partials: not available on a batch leader");
  const int64_t count = (int64_t)c->p.h_nsplit * c->p.h_ncb;
  if (capacity < count) return fail(NBMF_ERR_ARG, "nbmf_loglik_partials: buffer too small");
  CUDA_TRY(cudaMemcpyAsync(out_host, c->at<double>(c->p.oLLpart), (size_t)count * 8, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  if (col_blocks) *col_blocks = c->p.h_ncb;
  if (row_splits) *row_splits = c->p.h_nsplit;
  return NBMF_OK;
}

// ------------------------------------------------------------------------------------ fit loop
extern "C" int nbmf_fit_begin(nbmf_ctx* c, int32_t max_iter, double tol) {
  int rc = require_data(c);
  if (rc) return rc;
  if (max_iter < 1) return fail(NBMF_ERR_ARG, "max_iter must be >= 1");
  if (max_iter > c->cfg.max_iter_cap) return fail(NBMF_ERR_ARG, "max_iter exceeds cfg.max_iter_cap");
  if (c->max_iter != max_iter || c->tol != tol) graph_drop(c);      // both are baked into the captured finalize launches
  c->max_iter = max_iter;
  c->tol = tol;
  c->enqueued = 0;
  c->tail_enqueued = false;
  c->poll_pending = false;
  reset_state_kernel<<<1, 1, 0, c->st>>>(c->state(), c->cfg.alpha, c->cfg.beta);
  // prior sums of the initial H are not needed for any recorded loss, but keep the buffer defined
  launch_prior_sums(c->cfg.dtype, c->H(), c->cfg.n, c->cfg.k, c->p.pl.kp, c->p.ldh, c->cfg.eps,
                    c->at<double>(c->p.oPrior), c->st);
  CHECK_LAUNCH(2);
  return NBMF_OK;
}


static int enqueue_iteration(nbmf_ctx* c) {
  int rc;
  // H pass on (W_t, H_t): partial C, D and the log-likelihood of iteration t-1's factors
  if ((rc = enqueue_h_pass(c, 1, true))) return rc;  // + loss_{t-1}, stop rule; may set done
  if ((rc = enqueue_h_epilogue(c))) return rc;      // H_{t+1}
  return enqueue_w_step(c);                         // W_{t+1} from H_{t+1}
}

// Small problems are bound by launch overhead (an iteration is ~7 kernels of a few microseconds each, and n_init /
// grid sweeps run dozens of such fits at once): kGraphIters iterations are captured once per fit and replayed.  The
// stop rule lives on the device (`done` turns every kernel into a no-op), so a replay never needs the host.
constexpr int kGraphIters = 4;
static bool graph_eligible(const nbmf_ctx* c) {
  return c->graph_state >= 0 && c->world == 1 && !c->profile && (double)c->cfg.m * (double)c->cfg.n <= (double)(1 << 24);
}
static int graph_build(nbmf_ctx* c) {
  if (cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();                              // e.g. the legacy default stream cannot be captured
    c->graph_state = -1;
    return NBMF_OK;
  }
  const long long before = g_launches.load();
  int rc = NBMF_OK;
  for (int i = 0; i < kGraphIters && !rc; ++i) rc = enqueue_iteration(c);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c->st, &g);
  c->graph_launches = g_launches.load() - before;
  g_launches -= c->graph_launches;                   // counted per replay instead
  if (rc || e != cudaSuccess || !g || cudaGraphInstantiate(&c->graph_exec, g, 0) != cudaSuccess) {
    cudaGetLastError();
    c->graph_exec = nullptr;
    c->graph_state = -1;
  } else {
    c->graph_state = 1;
  }
  if (g) cudaGraphDestroy(g);
  return rc;
}

// Small fits: whole iterations inside one persistent cooperative launch (fused_small.cu) instead of ~6 launches each.
// A batch of fits (nbmf_batch_bind) shares the launch, a few CTAs per fit; batches larger than the co-resident grid run in
// slices.  The results do not depend on the grid (see the kernel), so batched and single fits stay bit-identical.
static bool fused_ok(const nbmf_ctx* c) { return c->p.fused && c->world == 1 && !c->profile; }

static int enqueue_fused(nbmf_ctx* c, int n_passes) {
  const Plan& p = c->p;
  int rc = eager_rowcount(c);
  if (rc) return rc;
  FusedArgs a;
  a.W = c->W(); a.H = c->H(); a.Ht = c->Ht(); a.P = c->P; a.M = c->M;
  a.m = c->cfg.m; a.n = c->cfg.n; a.ldh = p.ldh; a.wpr = p.wpr;
  a.k = c->cfg.k; a.kp = p.pl.kp; a.strict = p.strict ? 1 : 0; a.projection = c->cfg.projection;
  a.rows_per_warp = p.f_rows_per_warp; a.nsuper = p.f_nsuper; a.nwords = p.f_nwords;
  a.CDpart = c->ws + p.oFCD; a.LLpart = c->at<double>(p.oFLL);
  a.prior_part = c->at<double>(p.oPrior); a.prior_part2 = c->at<double>(p.oFPrior2); a.n_prior = p.n_prior;
  a.wsplit = p.f_wsplit;
  a.state = c->state(); a.history = c->at<double>(p.oHist);
  a.rowcount = (c->cfg.projection == NBMF_PROJ_DUCHI && c->M) ? (const void*)(c->ws + p.oRowcount) : nullptr;
  a.eps = c->cfg.eps; a.n_obs = c->cfg.n_obs; a.tol = c->tol; a.max_iter = c->max_iter;
  a.bar = c->at<unsigned>(p.oBar); a.n_passes = n_passes; a.h_in_smem = p.f_h_in_smem;
  a.batch_stride = c->batch_n > 1 ? c->batch_stride : 0;
  if (!c->fused_trace && getenv("NBMF_FUSED_TRACE")) {
    if (cudaMalloc((void**)&c->fused_trace, 4096 * 16 * 8) == cudaSuccess) cudaMemsetAsync(c->fused_trace, 0, 4096 * 16 * 8, c->st);
    else c->fused_trace = nullptr;
  }
  a.trace = c->fused_trace;
  if (c->fused_max_blocks < 0) c->fused_max_blocks = fused_max_blocks(c->cfg.dtype, a, p.f_smem);
  if (c->fused_max_blocks < 1) return fail(NBMF_ERR_CUDA, "fused small-fit kernel: no co-resident grid on this device");
  const int want = std::max(std::max(p.f_nwords * p.f_nsuper, (int)((c->cfg.m + 7) / 8)), p.n_prior);
  for (int f0 = 0; f0 < c->batch_n; f0 += c->fused_max_blocks) {
    const int nf = std::min(c->batch_n - f0, c->fused_max_blocks);
    const int grid_x = std::max(1, std::min(want, c->fused_max_blocks / nf));
    FusedArgs b = a;
    const size_t sh = (size_t)f0 * (size_t)a.batch_stride;
    b.W = (unsigned char*)a.W + sh; b.H = (unsigned char*)a.H + sh; b.Ht = (unsigned char*)a.Ht + sh;
    b.CDpart = (unsigned char*)a.CDpart + sh; b.LLpart = (double*)((unsigned char*)a.LLpart + sh);
    b.prior_part = (double*)((unsigned char*)a.prior_part + sh); b.prior_part2 = (double*)((unsigned char*)a.prior_part2 + sh);
    b.state = (FitState*)((unsigned char*)a.state + sh);
    b.history = (double*)((unsigned char*)a.history + sh); b.bar = (unsigned*)((unsigned char*)a.bar + sh);
    CUDA_TRY(cudaMemset2DAsync(b.bar, nf > 1 ? (size_t)a.batch_stride : 64, 0, 4, (size_t)nf, c->st));
    const int e = launch_fused_fit(c->cfg.dtype, b, grid_x, nf, p.f_smem, c->st);
    if (e) return cuda_fail((cudaError_t)e, "cudaLaunchCooperativeKernel(fused_fit_kernel)");
    g_launches += 1;
  }
  return NBMF_OK;
}

extern "C" int nbmf_fit_enqueue(nbmf_ctx* c, int32_t n_iters) {
  if (!c || c->max_iter < 1) return fail(NBMF_ERR_ARG, "nbmf_fit_begin was not called");
  int rc;
  if (fused_ok(c)) {
    const int iters = std::max(0, std::min((int)n_iters, c->max_iter - c->enqueued));
    const bool tail = c->enqueued + iters >= c->max_iter && !c->tail_enqueued;   // + the loss-only pass after the last iteration
    if (iters + (tail ? 1 : 0) > 0 && (rc = enqueue_fused(c, iters + (tail ? 1 : 0)))) return rc;
    c->enqueued += iters;
    if (tail) c->tail_enqueued = true;
    return NBMF_OK;
  }
  for (int i = 0; i < n_iters && c->enqueued < c->max_iter;) {
    // the first iteration always runs uncaptured (lazy one-time setup: kernel attributes, row counts)
    if (graph_eligible(c) && c->enqueued >= 1 && n_iters - i >= kGraphIters && c->max_iter - c->enqueued >= kGraphIters) {
      if (c->graph_state == 0 && (rc = graph_build(c))) return rc;
      if (c->graph_state == 1) {
        CUDA_TRY(cudaGraphLaunch(c->graph_exec, c->st));
        g_launches += c->graph_launches;
        c->enqueued += kGraphIters;
        i += kGraphIters;
        continue;
      }
    }
    if ((rc = enqueue_iteration(c))) return rc;
    c->enqueued += 1;
    ++i;
  }
  if (c->enqueued >= c->max_iter && !c->tail_enqueued) {
    // loss of the last iteration: one loss-only pass
    if ((rc = enqueue_h_pass(c, 0, true))) return rc;
    c->tail_enqueued = true;
  }
  return NBMF_OK;
}

extern "C" int nbmf_fit_poll(nbmf_ctx* c, int wait, int32_t* done_host, int32_t* n_iter_host) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  if (!c->poll_pending) {
    CUDA_TRY(cudaMemcpyAsync(c->host_state, c->state(), sizeof(FitState), cudaMemcpyDeviceToHost, c->st));
    CUDA_TRY(cudaEventRecord(c->poll_ev, c->st));
    c->poll_pending = true;
  }
  if (wait) {
    CUDA_TRY(cudaEventSynchronize(c->poll_ev));
  } else {
    cudaError_t e = cudaEventQuery(c->poll_ev);
    if (e == cudaErrorNotReady) {
      if (done_host) *done_host = -1;            // snapshot not ready yet
      return NBMF_OK;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventQuery");
  }
  c->poll_pending = false;
  if (done_host) *done_host = c->host_state->done;
  if (n_iter_host) *n_iter_host = c->host_state->n_hist;
  return NBMF_OK;
}

// ---- batched small fits (restarts, n_init): one launch per phase advances all fits of a group.
// The caller creates n contexts of identical configuration (same m, n, k, dtype, eps, flags, data planes; alpha and beta
// may differ from fit to fit: they live in each fit's FitState, so hyper-parameter grids batch like restarts)
// on the same stream, with workspaces at a uniform byte stride inside one allocation (leader first), gives each its
// factors (nbmf_set_factors) and calls nbmf_fit_begin on each with the same max_iter / tol.  After nbmf_batch_bind
// the leader's nbmf_fit_enqueue drives all of them: every kernel of the loop runs with gridDim.z = n and shifts its
// workspace pointers by blockIdx.z * stride; each fit keeps its own device-side state (loss history, stop rule,
// done flag), so fits that converge early simply turn into no-ops.  Both engines, single GPU.
extern "C" int nbmf_batch_bind(nbmf_ctx* c, int32_t n, int64_t stride_bytes) {
  if (!c || n < 1 || (n > 1 && stride_bytes < (int64_t)c->p.total) || (stride_bytes % 16) != 0 || n > 65535)
    return fail(NBMF_ERR_ARG, "nbmf_batch_bind: bad arguments");
  if (n > 1 && c->world != 1)
    return fail(NBMF_ERR_ARG, "nbmf_batch_bind: batches run on a single GPU");
  c->batch_n = n;
  c->batch_stride = n > 1 ? stride_bytes : 0;
  graph_drop(c);
  return NBMF_OK;
}
// Tail of every fit of the leader's batch with ONE synchronisation (per-fit nbmf_fit_history + nbmf_simplex_deviation cost
// two each: 7 ms for 64 fits): history_host[i * hist_stride ..] = the losses of fit i (hist_stride <= max_iter_cap + 2
// entries are copied per fit; the caller knows n_iter from nbmf_batch_poll), converged_host[i], deviation_host[i] =
// max |row sum of W - 1| of fit i (NaN if a row sum is not finite: the test of _solver.py:195-199).
extern "C" int nbmf_batch_tail(nbmf_ctx* c, double* history_host, int32_t hist_stride, int32_t* converged_host,
                               double* deviation_host) {
  if (!c || !history_host || !converged_host || !deviation_host || hist_stride < 1)
    return fail(NBMF_ERR_ARG, "nbmf_batch_tail: bad arguments");
  const int n = c->batch_n;
  const size_t pitch = n > 1 ? (size_t)c->batch_stride : c->p.total;
  const int count = std::min<int>(hist_stride, std::max(c->cfg.max_iter_cap, 1) + 2);
  for (int i = 0; i < n; ++i) {                        // deviation of fit i into its own scratch (64 bytes at oLoss)
    unsigned char* wsi = c->ws + (size_t)i * pitch;
    launch_simplex_deviation(c->cfg.dtype, wsi + c->p.oW, c->cfg.m, c->cfg.k, c->p.pl.kp,
                             reinterpret_cast<unsigned long long*>(wsi + c->p.oLoss), c->st);
    CHECK_LAUNCH(1);
  }
  std::vector<FitState> st((size_t)n);
  std::vector<unsigned long long> dv((size_t)n * 2);
  CUDA_TRY(cudaMemcpy2DAsync(history_host, (size_t)hist_stride * 8, c->at<double>(c->p.oHist), pitch, (size_t)count * 8, (size_t)n,
                             cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaMemcpy2DAsync(st.data(), sizeof(FitState), c->state(), pitch, sizeof(FitState), (size_t)n, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaMemcpy2DAsync(dv.data(), 16, c->ws + c->p.oLoss, pitch, 16, (size_t)n, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  for (int i = 0; i < n; ++i) {
    converged_host[i] = st[(size_t)i].converged;
    double d;
    memcpy(&d, &dv[(size_t)2 * i], 8);
    deviation_host[i] = dv[(size_t)2 * i + 1] ? NAN : d;
  }
  return NBMF_OK;
}
// states of all fits of the leader's batch: *all_done = every fit has stopped, n_iter_host[i] = losses recorded by fit i
extern "C" int nbmf_batch_poll(nbmf_ctx* c, int32_t* all_done, int32_t* n_iter_host) {
  if (!c || !all_done) return fail(NBMF_ERR_ARG, "nbmf_batch_poll: null argument");
  std::vector<FitState> h((size_t)c->batch_n);
  const size_t pitch = c->batch_n > 1 ? (size_t)c->batch_stride : sizeof(FitState);
  CUDA_TRY(cudaMemcpy2DAsync(h.data(), sizeof(FitState), c->state(), pitch, sizeof(FitState), (size_t)c->batch_n,
                             cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  int done = 1;
  for (int i = 0; i < c->batch_n; ++i) {
    done &= h[(size_t)i].done != 0;
    if (n_iter_host) n_iter_host[i] = h[(size_t)i].n_hist;
  }
  *all_done = done;
  return NBMF_OK;
}

extern "C" int nbmf_fit_history(nbmf_ctx* c, double* history_host, int32_t count, int32_t* converged_host) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  if (count > 0 && history_host)
    CUDA_TRY(cudaMemcpyAsync(history_host, c->at<double>(c->p.oHist), (size_t)count * 8, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaMemcpyAsync(c->host_state, c->state(), sizeof(FitState), cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  if (converged_host) *converged_host = c->host_state->converged;
  return NBMF_OK;
}

extern "C" int nbmf_fit(nbmf_ctx* c, int32_t max_iter, double tol, double* history_host, int32_t* n_iter_host,
                        int32_t* converged_host) {
  int rc = nbmf_fit_begin(c, max_iter, tol);
  if (rc) return rc;
  int chunk = 8;
  int32_t done = 0, n_iter = 0;
  // keep one chunk in flight beyond the one being waited on; converged iterations are no-ops
  if ((rc = nbmf_fit_enqueue(c, chunk))) return rc;
  while (true) {
    CUDA_TRY(cudaMemcpyAsync(c->host_state, c->state(), sizeof(FitState), cudaMemcpyDeviceToHost, c->st));
    CUDA_TRY(cudaEventRecord(c->poll_ev, c->st));
    const bool all_enqueued = c->tail_enqueued;
    if (!all_enqueued) {
      chunk = std::min(chunk * 2, 64);
      if ((rc = nbmf_fit_enqueue(c, chunk))) return rc;
    }
    CUDA_TRY(cudaEventSynchronize(c->poll_ev));
    done = c->host_state->done;
    n_iter = c->host_state->n_hist;
    if (done || all_enqueued) break;
  }
  CUDA_TRY(cudaMemcpyAsync(c->host_state, c->state(), sizeof(FitState), cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  n_iter = c->host_state->n_hist;
  if (n_iter_host) *n_iter_host = n_iter;
  return nbmf_fit_history(c, history_host, n_iter, converged_host);
}

// ------------------------------------------------------------------------------------ transform
extern "C" int nbmf_transform(nbmf_ctx* c, int32_t n_steps) {
  int rc = require_data(c);
  if (rc) return rc;
  for (int i = 0; i < n_steps; ++i)
    if ((rc = enqueue_w_step(c))) return rc;
  launch_clip_rows(c->cfg.dtype, c->W(), c->cfg.m, c->cfg.k, c->p.pl.kp, 1e-8, 1.0, c->st);
  CHECK_LAUNCH(1);
  return format_w(c, false);
}

// ------------------------------------------------------------------------------------ comm
extern "C" int nbmf_comm_unique_id(void* id128) {
  if (!id128) return fail(NBMF_ERR_ARG, "null id buffer");
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  rc = g_nccl.GetUniqueId(&id);
  if (rc) return nccl_fail(rc, "ncclGetUniqueId");
  memcpy(id128, &id, 128);
  return NBMF_OK;
}
extern "C" int nbmf_comm_init(nbmf_ctx* c, const void* id128, int32_t rank, int32_t world) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(NBMF_ERR_ARG, "nbmf_comm_init: bad arguments");
  if (world == 1) { c->world = 1; c->rank = 0; return NBMF_OK; }
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  memcpy(&id, id128, 128);
  rc = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (rc) return nccl_fail(rc, "ncclCommInitRank");
  c->owns_comm = true;
  c->world = world;
  c->rank = rank;
  return NBMF_OK;
}
// A communicator that outlives contexts: ncclCommInitRank costs ~1 s, a caller that fits many problems on
// the same ranks creates it once and attaches it to every context.
extern "C" int nbmf_comm_create(const void* id128, int32_t rank, int32_t world, void** comm_out) {
  if (!id128 || !comm_out || world < 2 || rank < 0 || rank >= world) return fail(NBMF_ERR_ARG, "nbmf_comm_create: bad arguments");
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  memcpy(&id, id128, 128);
  void* comm = nullptr;
  rc = g_nccl.CommInitRank(&comm, world, id, rank);
  if (rc) return nccl_fail(rc, "ncclCommInitRank");
  *comm_out = comm;
  return NBMF_OK;
}
extern "C" int nbmf_comm_attach(nbmf_ctx* c, void* comm, int32_t rank, int32_t world) {
  if (!c || !comm || world < 2 || rank < 0 || rank >= world) return fail(NBMF_ERR_ARG, "nbmf_comm_attach: bad arguments");
  if (c->comm && c->owns_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  c->comm = comm;
  c->owns_comm = false;
  c->world = world;
  c->rank = rank;
  graph_drop(c);
  return NBMF_OK;
}
extern "C" int nbmf_comm_destroy(void* comm) {
  if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
  return NBMF_OK;
}
extern "C" int nbmf_comm_world(nbmf_ctx* c) { return c ? c->world : 0; }
extern "C" int nbmf_engine(nbmf_ctx* c) {
  return !c ? 0 : (c->p.tensor ? NBMF_ENGINE_TENSOR : (c->p.fused ? NBMF_ENGINE_FUSED : NBMF_ENGINE_SIMT));
}
extern "C" int nbmf_fit_is_fused(nbmf_ctx* c) { return c && fused_ok(c) ? 1 : 0; }

// ------------------------------------------------------------------------------------ measurement
static void prof_clear(std::vector<cudaEvent_t>& v) {
  for (cudaEvent_t e : v) cudaEventDestroy(e);
  v.clear();
}
static void prof_sum(std::vector<cudaEvent_t>& v, double* ms, int32_t* count) {
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < v.size(); i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, v[i], v[i + 1]) == cudaSuccess) { tot += t; ++n; }
  }
  if (ms) *ms = tot;
  if (count) *count = n;
}
extern "C" int nbmf_profile_enable(nbmf_ctx* c, int enable) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  prof_clear(c->prof_h);
  prof_clear(c->prof_w);
  c->profile = enable != 0;
  graph_drop(c);
  return NBMF_OK;
}
extern "C" int nbmf_profile_read(nbmf_ctx* c, double* h_ms, int32_t* h_count, double* w_ms, int32_t* w_count) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  CUDA_TRY(cudaStreamSynchronize(c->st));
  prof_sum(c->prof_h, h_ms, h_count);
  prof_sum(c->prof_w, w_ms, w_count);
  return NBMF_OK;
}
extern "C" int nbmf_plan_info(nbmf_ctx* c, int32_t* h_col_blocks, int32_t* h_row_splits, int32_t* w_row_blocks,
                              int32_t* w_col_splits) {
  if (!c) return fail(NBMF_ERR_ARG, "null context");
  if (h_col_blocks) *h_col_blocks = c->p.h_ncb;
  if (h_row_splits) *h_row_splits = c->p.h_nsplit;
  if (w_row_blocks) *w_row_blocks = (int32_t)((c->cfg.m + c->p.pl.w_bmr - 1) / c->p.pl.w_bmr);
  if (w_col_splits) *w_col_splits = c->p.w_nsplit;
  return NBMF_OK;
}
extern "C" int nbmf_fma_peak(int dtype, int32_t iters, void* scratch, void* stream, double* tflops_host) {
  if (!scratch || !tflops_host || iters < 1) return fail(NBMF_ERR_ARG, "nbmf_fma_peak: bad arguments");
  *tflops_host = run_fma_peak(dtype, iters, (cudaStream_t)stream, (float*)scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "nbmf_fma_peak");
  return NBMF_OK;
}

// Whole MM iterations of a SMALL fit in one persistent kernel (reference: src/nbmf_mm/_solver.py:19-59 for the update,
// :148-175 for the loss and the stop rule).
//
// A 1226 x 285 fit (the paper's lastfm data) is a few hundred thousand entries: the pass kernels of the regular path finish
// in a few microseconds each and an iteration is bound by its six dependent launches (60-80 us).  Here a co-resident grid
// runs the iterations itself, with three grid-wide barriers per iteration instead of launches:
//
//   phase H   CTA unit = (32-column word, block of 8 R rows): every warp streams R rows of W against the word's column
//             tile of H (Theta, masked ratios, the two contractions, the fused log-likelihood), the eight warps' sums are
//             added in warp order and stored as the unit's partial C | D.
//   -- barrier --  every CTA then takes the same decision about the loss / stop rule from the same partials, in the same
//             order (finalize_core; only CTA 0 writes the history and the state).
//   phase H'  the H epilogue of misc_kernels.cu over virtual blocks (same prior partial layout), summing the row blocks'
//             partials in fixed order.
//   -- barrier --
//   phase W   one warp per row: Theta against H (staged in shared memory when it fits), the gradient reduced across the
//             warp, then the row's multiplicative step and projection (w_row_project) in the same warp.
//   -- barrier --
//
// Units, partial layouts and every summation order depend on the problem's shape only, never on the grid: results are
// bit-identical whatever the number of CTAs -- a batch of fits (blockIdx.y, a few CTAs each) equals the same fits run one
// at a time.  Data written inside the launch is read back with ld.global.cg (L2): L1 is not coherent across SMs.
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "fused_args.h"
#include "internal.h"

namespace nbmf {

namespace {

constexpr int FNT = 256;            // threads per CTA
constexpr int FNW = FNT / 32;       // warps per CTA

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblk, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblk;
    // release: this CTA's stores (ordered before by the bar.sync above) become visible before the arrival
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

template <typename Real>
__device__ __forceinline__ Real ldcg_(const Real* p) { return __ldcg(p); }

// Sums of KP per-lane values across the warp, KP - 1 shuffles (+ log2(32 / KP) for the lanes that end up with the same
// component) instead of 5 KP: in the round with lane offset o >= KP / 2 ... 1 a lane keeps the half of its values whose
// component has bit o equal to its own lane bit and adds the partner's partials for them; the groups of KP lanes are added
// last.  Lane l returns the total of component l % KP; fixed association, hence deterministic.
template <typename Real, int KP>
__device__ __forceinline__ Real warp_sum_transposed(Real (&g)[KP], int lane) {
#pragma unroll
  for (int o = (KP > 16 ? 16 : KP / 2); o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const Real keep = upper ? g[i + o] : g[i];
      const Real send = upper ? g[i] : g[i + o];
      g[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  Real t = g[0];                                       // total of component lane % KP over this lane's aligned group of KP lanes
#pragma unroll
  for (int o = KP; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

constexpr int kTraceRows = 4096;
__device__ __forceinline__ void stamp(const FusedArgs& a, int it, int slot) {
  if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && it < kTraceRows) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[it * 16 + slot] = t;
  }
}

template <typename Real, int KP>
__global__ void __launch_bounds__(FNT, 1) fused_fit_kernel(const FusedArgs a0) {
  // ---- this fit's workspace (blockIdx.y of a batch)
  const size_t sh = (size_t)blockIdx.y * (size_t)a0.batch_stride;
  Real* __restrict__ W = batch_shift(reinterpret_cast<Real*>(a0.W), sh);
  Real* __restrict__ H = batch_shift(reinterpret_cast<Real*>(a0.H), sh);
  Real* __restrict__ Ht = batch_shift(reinterpret_cast<Real*>(a0.Ht), sh);
  Real* __restrict__ CDp = batch_shift(reinterpret_cast<Real*>(a0.CDpart), sh);
  double* __restrict__ LLp = batch_shift(a0.LLpart, sh);
  double* const prior_b0 = batch_shift(a0.prior_part, sh);
  double* const prior_b1 = batch_shift(a0.prior_part2, sh);
  FitState* state = batch_shift(a0.state, sh);
  double* history = batch_shift(a0.history, sh);
  unsigned* bar = batch_shift(a0.bar, sh);
  const uint32_t* __restrict__ P = a0.P;
  const uint32_t* __restrict__ M = a0.M;
  const Real* __restrict__ rowcount = reinterpret_cast<const Real*>(a0.rowcount);

  const int64_t m = a0.m, n = a0.n, ldh = a0.ldh, wpr = a0.wpr;
  const int k = a0.k, kp = a0.kp, R = a0.rows_per_warp, nsuper = a0.nsuper, nwords = a0.nwords;
  const int64_t ldw = (int64_t)nwords * 32;           // leading dimension of the partials and of the staged H
  const int nunits = nwords * nsuper;
  const bool strict = a0.strict != 0, has_mask = (M != nullptr);
  const Real eps = (Real)a0.eps;
  const unsigned nblk = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too: branches on it stay converged

  extern __shared__ __align__(16) unsigned char smem_raw[];
  Real* const smem = reinterpret_cast<Real*>(smem_raw);
  __shared__ FitState ls;
  __shared__ double red[FNW], red2[FNW];
  __shared__ double sG[FNW][33];                       // W phase: partial gradients of the warps that share a row

  if (tid == 0) ls = *state;
  __syncthreads();
  unsigned target = 0;

  for (int pass = 0; pass < a0.n_passes; ++pass) {
    if (ls.done) break;                               // the same on every CTA
    const bool cd = ls.it < a0.max_iter;              // else: the loss-only pass after the last iteration
    const int trace_it = ls.it;
    stamp(a0, trace_it, 0);

    // =============================================================== phase H
    {
      Real* const sH = smem;                          // [KP][32]   the word's column tile of H
      Real* const sW = sH + KP * 32;                  // [FNW][R][KP] rows of W, per warp
      Real* const sAcc = sW + FNW * R * KP;           // [FNW][2][KP][32] the warps' C | D
      Real* const sWw = sW + warp * R * KP;
      #pragma unroll 1
      for (int u = blockIdx.x; u < nunits; u += nblk) {
        const int wd = u / nsuper, sc = u % nsuper;
        #pragma unroll 1
        for (int e = tid; e < k * 32; e += FNT) sH[e] = ldcg_(H + (int64_t)(e >> 5) * ldh + 32 * wd + (e & 31));
        const int64_t r0 = (int64_t)sc * (FNW * R) + (int64_t)warp * R;
        const int nrows = (int)max((int64_t)0, min((int64_t)R, m - r0));
        #pragma unroll 1
        for (int e = lane; e < nrows * k; e += 32) {
          const int r = e / k, kk = e - r * k;
          sWw[r * KP + kk] = ldcg_(W + (r0 + r) * kp + kk);
        }
        // bit words of this warp's rows: lane r holds row r0 + r
        uint32_t pwl = 0u, mwl = 0xffffffffu;
        if (lane < nrows) {
          pwl = __ldg(P + (r0 + lane) * wpr + wd);
          if (strict) mwl = __ldg(M + (r0 + lane) * wpr + wd);
        }
        __syncthreads();
        if (u == (int)blockIdx.x) stamp(a0, trace_it, 8);

        Real c[KP], d[KP];
#pragma unroll
        for (int kk = 0; kk < KP; ++kk) c[kk] = d[kk] = Real(0);
        // log of a product of LOGG factors instead of LOGG logs (x >= eps: fp32 needs eps >= 1e-9, checked by the plan);
        // a negative factor is flagged so that the result is NaN, as the reference's log(x) would be
        constexpr int LOGG = sizeof(Real) == 8 ? 8 : 4;
        Real px = Real(1), ll = Real(0);
        bool neg = false;
        double lld = 0.0;
        #pragma unroll 1                              // (two rows in flight: measured slower, the register file is full)
        for (int r = 0; r < nrows; ++r) {
          const Real* wr = sWw + r * KP;
          Real theta = Real(0);
#pragma unroll
          for (int kk = 0; kk < KP; ++kk)
            if (kk < k) theta = fma(wr[kk], sH[kk * 32 + lane], theta);
          const bool p = (__shfl_sync(0xffffffffu, pwl, r) >> lane) & 1u;
          const Real x = (p ? theta : (Real(1) - theta)) + eps;
          Real r_ = rcp_(x);
          Real xf = x;
          if (strict) {
            const bool o = (__shfl_sync(0xffffffffu, mwl, r) >> lane) & 1u;
            r_ = o ? r_ : Real(0);
            xf = o ? x : Real(1);
          }
          px *= xf;
          neg |= (xf < Real(0));
          if ((r & (LOGG - 1)) == LOGG - 1 || r == nrows - 1) {
            ll += logu_(neg ? Real(NAN) : px);
            px = Real(1);
            neg = false;
          }
          if ((r & 63) == 63) { lld += (double)ll; ll = Real(0); }
          if (cd) {
            const Real rp = p ? r_ : Real(0), rn = p ? Real(0) : r_;
#pragma unroll
            for (int kk = 0; kk < KP; ++kk)
              if (kk < k) {
                c[kk] = fma(wr[kk], rp, c[kk]);
                d[kk] = fma(wr[kk], rn, d[kk]);
              }
          }
        }
        if (u == (int)blockIdx.x) stamp(a0, trace_it, 9);
        lld += (double)ll;
        if (32 * (int64_t)wd + lane >= n) lld = 0.0;  // padded columns
        lld = warp_sum(lld);
        if (lane == 0) red[warp] = lld;
        // the eight warps' sums: every warp leaves its own in shared memory, then each element is added in warp order
        if (cd) {
          Real* const mine = sAcc + warp * (2 * KP * 32);
#pragma unroll
          for (int kk = 0; kk < KP; ++kk)
            if (kk < k) {
              mine[kk * 32 + lane] = c[kk];
              mine[KP * 32 + kk * 32 + lane] = d[kk];
            }
        }
        __syncthreads();
        if (cd) {
#pragma unroll 1
          for (int e = tid; e < 2 * k * 32; e += FNT) {
            const int which = e / (k * 32), rem = e - which * (k * 32);    // rem = kk * 32 + lane
            const Real* src = sAcc + which * KP * 32 + rem;
            Real t = src[0];
#pragma unroll
            for (int w8 = 1; w8 < FNW; ++w8) t += src[w8 * (2 * KP * 32)];
            CDp[((int64_t)(sc * 2 + which) * kp + (rem >> 5)) * ldw + 32 * wd + (rem & 31)] = t;
          }
        }
        if (u == (int)blockIdx.x) stamp(a0, trace_it, 10);
        if (tid == 0) {
          double t = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < FNW; ++w8) t += red[w8];
          LLp[u] = t * log_unit<Real>();
        }
        __syncthreads();                              // sH, sW, sAcc and red are reused by the next unit
      }
    }
    stamp(a0, trace_it, 1);
    grid_barrier(bar, nblk, target);
    stamp(a0, trace_it, 2);

    // ---- loss of the previous iteration + stop rule: every CTA, from the same numbers in the same order
    // (the prior partial sums of the current H are double-buffered by iteration parity: a CTA that is already in the H
    // epilogue below writes the other buffer while a slower CTA still reads this one; buffer 0 is the one nbmf_fit_begin
    // fills for the initial H)
    if (warp == 0) {
      const double* __restrict__ prior = (trace_it & 1) ? prior_b1 : prior_b0;
      double v = 0.0, pa = 0.0, pb = 0.0;
      #pragma unroll 8                                // independent L2 loads in flight, fixed order of the adds
      for (int i = lane; i < nunits; i += 32) v += __ldcg(LLp + i);
      #pragma unroll 4
      for (int i = lane; i < a0.n_prior; i += 32) {
        pa += __ldcg(prior + 2 * i);
        pb += __ldcg(prior + 2 * i + 1);
      }
      v = warp_sum(v);
      pa = warp_sum(pa);
      pb = warp_sum(pb);
      if (lane == 0) {
        finalize_core(ls, v, pa, pb, a0.n_obs, a0.tol, a0.max_iter, blockIdx.x == 0 ? history : nullptr);
        if (blockIdx.x == 0) *state = ls;
      }
    }
    __syncthreads();
    if (ls.done) break;
    stamp(a0, trace_it, 3);

    // =============================================================== phase H': epilogue over virtual blocks
    {
      const int nbx = (int)((n + FNT - 1) / FNT);
      const double alpha = ls.alpha, beta = ls.beta;
      double* __restrict__ prior = (trace_it & 1) ? prior_b0 : prior_b1;
      #pragma unroll 1
      for (int vb = blockIdx.x; vb < a0.n_prior; vb += nblk) {
        const int kk = vb / nbx;
        const int64_t j = (int64_t)(vb - kk * nbx) * FNT + tid;
        double la = 0.0, lb = 0.0;
        if (kk < k && j < n) {
          Real cs = ldcg_(CDp + (int64_t)kk * ldw + j), ds = ldcg_(CDp + ((int64_t)kp + kk) * ldw + j);
          #pragma unroll 8
          for (int s = 1; s < nsuper; ++s) {
            cs += ldcg_(CDp + ((int64_t)(s * 2) * kp + kk) * ldw + j);
            ds += ldcg_(CDp + ((int64_t)(s * 2 + 1) * kp + kk) * ldw + j);
          }
          const int64_t o = (int64_t)kk * ldh + j;
          const Real hn = h_update_elem<Real>(ldcg_(H + o), cs, ds, alpha, beta, a0.eps);
          H[o] = hn;
          Ht[j * kp + kk] = hn;
          la = log((double)(hn + eps));
          lb = log((double)((Real(1) - hn) + eps));
        }
        // both sums with one pair of barriers: fixed shuffle tree, then the warps in order (as block_sum does)
        la = warp_sum(la);
        lb = warp_sum(lb);
        __syncthreads();
        if (lane == 0) { red[warp] = la; red2[warp] = lb; }
        __syncthreads();
        if (tid == 0) {
          double sa = 0.0, sb = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < FNW; ++w8) { sa += red[w8]; sb += red2[w8]; }
          prior[2 * vb] = sa;
          prior[2 * vb + 1] = sb;
        }
      }
    }
    stamp(a0, trace_it, 4);
    grid_barrier(bar, nblk, target);
    stamp(a0, trace_it, 5);

    // =============================================================== phase W: `wsplit` warps per row, epilogue included
    {
      const bool staged = a0.h_in_smem != 0;
      const int ws = a0.wsplit, rows_per_cta = FNW / ws, sub = warp % ws, rloc = warp / ws;
      const int64_t ngroups = (m + rows_per_cta - 1) / rows_per_cta;
      if (staged && (int64_t)blockIdx.x < ngroups) {           // H (k x ldw) -> shared memory, through L2 (cp.async.cg)
        const int cpr = (int)(ldw * sizeof(Real) / 16);        // 16-byte chunks per row
        #pragma unroll 1
        for (int c = tid; c < k * cpr; c += FNT) {
          const int kk = c / cpr, q = c - kk * cpr;
          cp_async16(smem_raw + ((size_t)kk * cpr + q) * 16, reinterpret_cast<const unsigned char*>(H + (int64_t)kk * ldh) + (size_t)q * 16);
        }
        cp_async_commit();
        cp_async_wait<0>();
      }
      __syncthreads();
      stamp(a0, trace_it, 11);
      #pragma unroll 1
      for (int64_t grp = blockIdx.x; grp < ngroups; grp += nblk) {
        const int64_t row = grp * rows_per_cta + rloc;
        const bool active = row < m;
        Real gm = Real(0), wm = Real(0), q = Real(0);           // component `lane` of this row (this warp's share of it)
        if (active) {
          Real w[KP], g[KP];
#pragma unroll
          for (int kk = 0; kk < KP; ++kk) {
            w[kk] = (kk < k) ? ldcg_(W + row * kp + kk) : Real(0);
            g[kk] = Real(0);
          }
          #pragma unroll 1
          for (int wb = 0; wb < nwords; wb += 32) {     // bit words of the row: lane l holds word wb + l
            uint32_t pwl = 0u, mwl = 0xffffffffu;
            if (wb + lane < nwords) {
              pwl = __ldg(P + row * wpr + wb + lane);
              if (has_mask) mwl = __ldg(M + row * wpr + wb + lane);
            }
            const int nw = min(32, nwords - wb);
            #pragma unroll 1
            for (int wi = sub; wi < nw; wi += ws) {     // wb is a multiple of 32 >= ws: word wb + wi belongs to warp wi % ws
              const int64_t j = 32 * (int64_t)(wb + wi) + lane;
              const bool p = (__shfl_sync(0xffffffffu, pwl, wi) >> lane) & 1u;
              const bool ob = ((__shfl_sync(0xffffffffu, mwl, wi) >> lane) & 1u) && j < n;
              // H[kk][j] is read twice (Theta, then the gradient) rather than held: w, g and a third K-vector do not
              // fit the register file in fp64 at K = 32
              Real t0 = Real(0), t1 = Real(0);
              if (staged) {
                const Real* hcol = smem + j;
#pragma unroll
                for (int kk = 0; kk < KP; kk += 2) {
                  if (kk < k) t0 = fma(w[kk], hcol[(int64_t)kk * ldw], t0);
                  if (kk + 1 < k) t1 = fma(w[kk + 1], hcol[(int64_t)(kk + 1) * ldw], t1);
                }
              } else {
#pragma unroll
                for (int kk = 0; kk < KP; kk += 2) {
                  if (kk < k) t0 = fma(w[kk], ldcg_(H + (int64_t)kk * ldh + j), t0);
                  if (kk + 1 < k) t1 = fma(w[kk + 1], ldcg_(H + (int64_t)(kk + 1) * ldh + j), t1);
                }
              }
              const Real theta = t0 + t1;
              const Real x = (p ? theta : (Real(1) - theta)) + eps;
              Real r_ = rcp_(x);
              r_ = ob ? r_ : Real(0);
              const Real s = p ? r_ : -r_;
              q += p ? Real(0) : r_;
              if (staged) {
                const Real* hcol = smem + j;
#pragma unroll
                for (int kk = 0; kk < KP; ++kk)
                  if (kk < k) g[kk] = fma(hcol[(int64_t)kk * ldw], s, g[kk]);
              } else {
#pragma unroll
                for (int kk = 0; kk < KP; ++kk)
                  if (kk < k) g[kk] = fma(ldcg_(H + (int64_t)kk * ldh + j), s, g[kk]);
              }
            }
          }
          if (grp == (int64_t)blockIdx.x) stamp(a0, trace_it, 12);
          q = warp_sum(q);
          gm = warp_sum_transposed<Real, KP>(g, lane);
#pragma unroll
          for (int kk = 0; kk < KP; ++kk)
            if (lane == kk) wm = w[kk];
        }
        if (grp == (int64_t)blockIdx.x) stamp(a0, trace_it, 13);
        if (ws > 1) {                                  // the row's warps, added in warp order
          Real* sGr = reinterpret_cast<Real*>(&sG[0][0]);
          sGr[warp * 33 + lane] = gm;
          if (lane == 0) sGr[warp * 33 + 32] = q;
          __syncthreads();
          if (sub == 0) {
            gm = sGr[warp * 33 + lane];
            q = sGr[warp * 33 + 32];
            #pragma unroll 1
            for (int s2 = 1; s2 < ws; ++s2) {
              gm += sGr[(warp + s2) * 33 + lane];
              q += sGr[(warp + s2) * 33 + 32];
            }
          }
        }
        if (sub == 0 && active) {
          const Real denom = (a0.projection == 0) ? (Real)n : (rowcount ? rowcount[row] : (Real)n);
          Real v[1];
          v[0] = (lane < k) ? (wm * (gm + q)) / denom : Real(0);
          w_row_project<Real, 1>(v, v[0], k, lane, a0.projection, W + row * kp);
        }
        if (ws > 1) __syncthreads();
      }
    }
    stamp(a0, trace_it, 6);
    grid_barrier(bar, nblk, target);
    stamp(a0, trace_it, 7);
  }
}

template <typename Real, int KP>
int launch_one(const FusedArgs& a, int grid_x, int batch_n, size_t smem, cudaStream_t st, int* max_blocks) {
  static std::atomic<unsigned long long> attr_set{0};
  ensure_dynamic_smem(fused_fit_kernel<Real, KP>, 200 * 1024, attr_set);
  if (max_blocks) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_fit_kernel<Real, KP>, FNT, smem) != cudaSuccess) per_sm = 0;
    *max_blocks = per_sm * sms;
    return 0;
  }
  void* args[] = {(void*)&a};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)fused_fit_kernel<Real, KP>, dim3((unsigned)grid_x, (unsigned)batch_n, 1),
                                                    dim3(FNT, 1, 1), args, smem, st);
  return e == cudaSuccess ? 0 : (int)e;
}

template <typename Real>
int dispatch_kp(const FusedArgs& a, int grid_x, int batch_n, size_t smem, cudaStream_t st, int* max_blocks) {
  if (a.k <= 8) return launch_one<Real, 8>(a, grid_x, batch_n, smem, st, max_blocks);
  if (a.k <= 16) return launch_one<Real, 16>(a, grid_x, batch_n, smem, st, max_blocks);
  return launch_one<Real, 32>(a, grid_x, batch_n, smem, st, max_blocks);
}

}  // namespace

int fused_kp(int k) { return k <= 8 ? 8 : (k <= 16 ? 16 : 32); }

size_t fused_smem_bytes(int dtype, int k, int rows_per_warp, int nwords, int h_in_smem) {
  const size_t sz = dtype == 0 ? 4 : 8;
  const size_t KP = (size_t)fused_kp(k);
  const size_t phase_h = (KP * 32 + (size_t)FNW * rows_per_warp * KP + (size_t)FNW * 2 * KP * 32) * sz;
  const size_t phase_w = h_in_smem ? (size_t)k * nwords * 32 * sz : 0;
  return phase_h > phase_w ? phase_h : phase_w;
}

// CTAs of this kernel that can be resident at once on the current device (0: the launch cannot be made)
int fused_max_blocks(int dtype, const FusedArgs& a, size_t smem) {
  int mb = 0;
  if (dtype == 0) dispatch_kp<float>(a, 0, 0, smem, nullptr, &mb);
  else dispatch_kp<double>(a, 0, 0, smem, nullptr, &mb);
  return mb;
}

// returns 0 or the cudaError_t of the cooperative launch
int launch_fused_fit(int dtype, const FusedArgs& a, int grid_x, int batch_n, size_t smem, cudaStream_t st) {
  return dtype == 0 ? dispatch_kp<float>(a, grid_x, batch_n, smem, st, nullptr) : dispatch_kp<double>(a, grid_x, batch_n, smem, st, nullptr);
}

}  // namespace nbmf

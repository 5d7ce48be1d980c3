// Per-element / per-row epilogue arithmetic shared by the stand-alone epilogue kernels (misc_kernels.cu) and the fused
// small-fit kernel (fused_small.cu): one definition of the H ratio, the W projection and the loss / stop rule.
#pragma once
#include <math.h>

#include "common.cuh"
#include "internal.h"

namespace nbmf {

// ---- loss of the previous iteration + stop rule (_solver.py:158-175) on a fit's state; ll = log-likelihood of the
// factors the H pass just saw, pa / pb = sum log(H + eps), sum log((1 - H) + eps) of the same H.  history may be NULL
// (a redundant copy of the decision that must not write).
__device__ __forceinline__ void finalize_core(FitState& s, double ll, double pa, double pb, double n_obs, double tol,
                                              int max_iter, double* history) {
  const int it = s.it;
  int done = 0;
  if (it >= 1) {
    // alpha, beta from the fit's own state (per fit in a batch)
    const double loss = -(ll + (s.alpha - 1.0) * pa + (s.beta - 1.0) * pb) / n_obs;
    if (history) history[it - 1] = loss;
    s.n_hist = it;
    if (it >= 2) {
      const double prev = s.prev_loss;
      const double rel = fabs(prev - loss) / fabs(prev);
      if (rel < tol) { done = 1; s.converged = 1; }
    }
    s.prev_loss = loss;
    if (it >= max_iter) done = 1;
  }
  s.prior_a = pa;
  s.prior_b = pb;
  if (done) s.done = 1;
  else s.it = it + 1;
}

// ---- H <- (H*C + alpha-1) / (H*C + (1-H)*D + alpha+beta-2 + eps), clipped to [eps, 1-eps]   (_solver.py:42-47)
template <typename Real>
__device__ __forceinline__ Real h_update_elem(Real h, Real c, Real d, double alpha, double beta, double eps_d) {
  const Real eps = (Real)eps_d;
  const Real num = h * c + (Real)(alpha - 1.0);
  const Real den = (Real(1) - h) * d + (Real)(beta - 1.0);
  Real hn = num / (num + den + eps);
  hn = fmin(fmax(hn, eps), Real(1) - eps);
  return hn;
}

template <typename Real>
__device__ __forceinline__ Real warp_sum(Real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename Real>
__device__ __forceinline__ Real warp_scan_incl(Real v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const Real t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// ---- simplex projection of one row held across a warp (lane = component k; EPL = 1, 2 or 4 components per lane: lane,
// lane + 32, ..) and the store of the row.  v = the multiplicative step W*G/denom, part = this lane's share of its sum.
// projection 0 = "normalize" (L1 renormalisation, _solver.py:55-57), 1 = "duchi" (Euclidean projection, Duchi et al.
// 2008): descending bitonic sort of the row across the warp's registers, inclusive prefix sums in sorted order, rho = last
// index whose value exceeds the running threshold, then w = max(v - theta, 0).  Every reduction is a fixed shuffle tree.
template <typename Real, int EPL>
__device__ __forceinline__ void w_row_project(Real (&v)[EPL], Real part, int k, int lane, int projection, Real* __restrict__ Wrow) {
  static_assert(EPL == 1 || EPL == 2 || EPL == 4, "components per lane");
  if (projection == 0) {
    const Real sum = warp_sum(part);
#pragma unroll
    for (int e = 0; e < EPL; ++e)
      if (lane + 32 * e < k) Wrow[lane + 32 * e] = v[e] / sum;
    return;
  }
  // ---- Duchi: bitonic sort (descending) of the 32 * EPL values, element index = lane + 32 e; padding sorts last
  const Real NEG = -INFINITY;
  Real u[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) u[e] = (lane + 32 * e < k) ? v[e] : NEG;
  constexpr int NEL = 32 * EPL;
#pragma unroll
  for (int size = 2; size <= NEL; size <<= 1) {
#pragma unroll
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j >= 32) {                               // partner is another register of this lane: e ^ (j / 32)
        const int je = j >> 5;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          if ((e & je) == 0 && (e | je) < EPL) {
            const int f = e | je;                   // indices lane + 32 e (lower) and lane + 32 f
            const bool desc = (((32 * e) & size) == 0);   // direction of the merge: descending when (index & size) == 0
            const Real lo = fmin(u[e], u[f]), hi = fmax(u[e], u[f]);
            u[e] = desc ? hi : lo;
            u[f] = desc ? lo : hi;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const int idx = lane + 32 * e;
          const Real other = __shfl_xor_sync(0xffffffffu, u[e], j);
          const bool desc = ((idx & size) == 0);
          const bool lower = ((lane & j) == 0);     // this element is the lower index of the pair
          const Real mx = fmax(u[e], other), mn = fmin(u[e], other);
          u[e] = (lower == desc) ? mx : mn;
        }
      }
    }
  }
  // inclusive prefix sums in sorted order (element index = lane + 32 e), padding contributes nothing
  Real css[EPL];
  Real carry = Real(0);
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const Real x = (u[e] == NEG) ? Real(0) : u[e];
    css[e] = warp_scan_incl(x, lane) + carry;
    carry = __shfl_sync(0xffffffffu, css[e], 31);
  }
  // theta = (css_rho - 1) / (rho + 1) for the LAST index rho with u_rho - (css_rho - 1) / (rho + 1) > 0
  int best = -1;
  Real theta = Real(0);
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int idx = lane + 32 * e;
    const Real t = (css[e] - Real(1)) / (Real)(idx + 1);
    const bool ok = (u[e] != NEG) && (u[e] - t > Real(0));
    if (ok && idx > best) { best = idx; theta = t; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
    const Real ot = __shfl_xor_sync(0xffffffffu, theta, o);
    if (ob > best) { best = ob; theta = ot; }
  }
#pragma unroll
  for (int e = 0; e < EPL; ++e)
    if (lane + 32 * e < k) Wrow[lane + 32 * e] = fmax(v[e] - theta, Real(0));
}

}  // namespace nbmf

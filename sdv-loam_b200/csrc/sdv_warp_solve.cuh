// sdv_warp_solve.cuh — warp-cooperative, register-resident 8x8 pivoted LDLT solve for the device-resident LM loop.
//
// Replaces the `Hl.ldlt().solve(-b)` of CoarseTracker.cpp:724-748 inside track_cluster_kernel.  Lane r (r = lane & 7) owns
// row r of the symmetric system in 8 registers; pivot search, row exchange and broadcasts are width-8 shuffles; all column
// indices are compile-time constants (fully unrolled), so nothing touches local memory.  Same algorithm family as
// Eigen::LDLT (symmetric diagonal pivoting by largest |diagonal|, first index on ties; D^+ with the epsilon*max|d| tolerance),
// in right-looking form.
#pragma once
#include <cuda_runtime.h>

namespace sdv {

__device__ __forceinline__ double sel8(const double (&a)[8], int i) {
  double v = a[0];
#pragma unroll
  for (int j = 1; j < 8; j++) v = (i == j) ? a[j] : v;
  return v;
}
__device__ __forceinline__ void put8(double (&a)[8], int i, double v) {
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = (i == j) ? v : a[j];
}

// All 32 lanes must call.  a[] = row (lane&7) of the n x n system padded with identity to 8x8, rhs = its right-hand side.
// Returns x[(lane&7)] for the ORIGINAL variable (lane&7).
__device__ __forceinline__ double warp_ldlt_solve8(double (&a)[8], double rhs) {
  const unsigned full = 0xffffffffu;
  const int r = threadIdx.x & 7;
  int idx = r;                                             // which original variable this (permuted) row stands for
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // ---- pivot: largest |diag| among rows >= k, first index on ties
    double d = sel8(a, r);
    double v = (r >= k) ? fabs(d) : -1.0; int vi = r;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(full, v, o, 8); int oi = __shfl_xor_sync(full, vi, o, 8);
      bool take = (ov > v) || (ov == v && oi < vi);
      v = take ? ov : v; vi = take ? oi : vi;
    }
    const int piv = vi;
    if (piv != k) {                                        // uniform across the 8-lane group
      const int src = (r == k) ? piv : ((r == piv) ? k : r);
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] = __shfl_sync(full, a[j], src, 8);
      rhs = __shfl_sync(full, rhs, src, 8); idx = __shfl_sync(full, idx, src, 8);
      double ak = a[k], ap = sel8(a, piv);
      a[k] = ap; put8(a, piv, ak);
    }
    // ---- eliminate column k
    const double dk = __shfl_sync(full, a[k], k, 8);
    const double rk = __shfl_sync(full, rhs, k, 8);
    const bool valid = fabs(dk) > 0;
    double l = (valid ? a[k] / dk : a[k]);
#pragma unroll
    for (int j = k + 1; j < 8; j++) {
      const double akj = __shfl_sync(full, a[j], k, 8);
      if (r > k) a[j] -= l * akj;
      else if (r == k && valid) a[j] = a[j] / dk;         // row k keeps L^T: a[j] = L[j][k]
    }
    if (r > k) { rhs -= l * rk; a[k] = l; }
  }
  // ---- D^+ (Eigen: tolerance = max|d| * eps)
  double d = sel8(a, r);
  double dmax = fabs(d);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(full, dmax, o, 8));
  double tol = dmax * 2.220446049250313e-16; if (tol < 1.0/1.7976931348623157e308) tol = 1.0/1.7976931348623157e308;
  double y = (fabs(d) > tol) ? rhs / d : 0.0;
  // ---- L^T x = y
#pragma unroll
  for (int k = 7; k > 0; k--) {
    const double xk = __shfl_sync(full, y, k, 8);
    if (r < k) y -= a[k] * xk;
  }
  // ---- un-permute: variable idx has value y; route it to lane idx
  double out = 0.0;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const double ys = __shfl_sync(full, y, s, 8); const int is = __shfl_sync(full, idx, s, 8);
    out = (is == r) ? ys : out;
  }
  return out;
}


} // namespace sdv

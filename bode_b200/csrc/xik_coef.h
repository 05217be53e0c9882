// Even Chebyshev series of the xik factors, shared by the two prepare kernels (prepare.cu, prepare_tiles.cu).
#pragma once
#include <math.h>

#include "layout.h"

namespace bode {

// ---- even Chebyshev series of the xik factors (candidate-independent part of xik(x), reference _core.py:86-92) -----
// F_k(x) = sqrt(pi)/sqrt(2) * ell_k * (erf((1-x) c_k) - erf((0-x) c_k)) is symmetric about x = 1/2, so with u = 2x-1 it is
// a series in T_2j(u) = T_j(w), w = 2u^2 - 1.  One CTA per (sample, dimension): sample F at the 64 Gauss-Chebyshev nodes
// w_i = cos(theta_i) (u_i = cos(theta_i / 2) exactly), a_j = (2/64) sum_i F_i cos(j theta_i) with the cosines taken
// through cospi of an exactly reduced argument, cut the series where the coefficients drop under the rounding floor of
// this very sum (4e-16 a_0).  The EKLD kernel then spends 2 n + 6 operations per factor instead of two erf (~130).
// Runs inside the k_prepare launch on the blocks with blockIdx.x >= S (they land on the SMs the S factorisation CTAs leave
// idle): block S + s * ceil(d/4) + q handles dimensions 4q..4q+3 of sample s, 64 threads per dimension.
__device__ __forceinline__ void xik_coef_block(const double* __restrict__ hyp, int ndim, int d, char* ws, const Layout& L, int s, int k0,
                                               double* F /*[4][64]*/, double* A /*[4][64]*/) {
  const int kq = threadIdx.x >> 6, i = threadIdx.x & 63, k = k0 + kq;
  const bool on = k < d;
  if (on) {
    const double ell = hyp[(size_t)s * ndim + 1 + k];
    const double c = 1.0 / (1.4142135623730951 * ell);
    const double u = cospi(((double)i + 0.5) / (2.0 * kXcNodes));  // cos(theta_i / 2), theta_i = pi (i + 1/2) / 64
    const double x = 0.5 * (u + 1.0);
    F[kq * kXcNodes + i] = 1.2533141373155001 * ell * (erf((1.0 - x) * c) - erf((0.0 - x) * c));
  }
  __syncthreads();
  if (on) {
    double acc = 0.0;
    for (int m = 0; m < kXcNodes; m++) {
      const int r = (i * (2 * m + 1)) & (4 * kXcNodes - 1);  // j theta_m = pi r / 128 (mod 2 pi), j = this thread
      acc = fma(F[kq * kXcNodes + m], cospi((double)r / (2.0 * kXcNodes)), acc);
    }
    A[kq * kXcNodes + i] = acc * (2.0 / kXcNodes);
  }
  __syncthreads();
  if (!on) return;
  const double* Ak = A + kq * kXcNodes;
  double* out = reinterpret_cast<double*>(ws + L.off_xc) + (size_t)s * L.xc_doubles + (size_t)k * kXcStride;
  if (i == 0) {
    const double floor_ = 4e-16 * fabs(Ak[0]);
    int n = 0;
    bool ok = Ak[0] == Ak[0];
    for (int j = 0; j < kXcNodes; j++) {
      if (!(fabs(Ak[j]) <= floor_)) {       // (also true for NaN)
        if (j < kXcMax) n = j + 1;
        else ok = false;                    // not converged within kXcMax terms: erf path
      }
    }
    out[0] = ok ? (double)n : 0.0;
  }
  if (i < kXcMax) out[1 + i] = (i == 0) ? 0.5 * Ak[0] : Ak[i];
}


}  // namespace bode

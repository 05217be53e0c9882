// Branch-free FP64 evaluation of  variance * exp(x),  x <= 0,  shared by the prepare and scoring kernels so that
// K(X, X) and k(X, x) are computed by the same arithmetic.
//
// exp(x) = 2^(k/32) * exp(r),  k = rint(x * 32/ln2),  |r| <= ln2/64:  the 32-entry table holds
// variance * 2^(j/32) (filled per hyper-sample by the prepare kernel), exp(r) - 1 is a degree-6 Taylor polynomial
// (truncation 3.4e-18), and 2^(k>>5) is applied by adding to the exponent field.  11 FP64-pipe operations and no
// branches (libm's exp costs ~20 and carries a slow-path branch that serialises independent evaluations);
// measured error <= 1.04 ulp against expl over [-690, 0] (tools/).  For x < -693 the result saturates
// at ~1e-301 * variance instead of underflowing -- a covariance that small is zero for every purpose here.
#pragma once

namespace bode {

constexpr int kExpTab = 32;

__device__ __forceinline__ double rbf_exp(double x, const double* __restrict__ tab) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 46.166241308446828, SHIFT);
  const double kf = t - SHIFT;
  double r = fma(kf, -0x1.62e42fefa0000p-6, x);
  r = fma(kf, -5.1456092446553382e-14, r);
  const int k = __double2loint(t);
  double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  const double q = fma(r * r, p, r);
  const double T = tab[k & (kExpTab - 1)];
  const double res = fma(T, q, T);
  const int e = max(k >> 5, -1000);
  return __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res));
}

// sum_k (u[k] - y[k])^2 over DP dimensions, u in shared/global memory (16-byte aligned rows), y in registers
template <int DP>
__device__ __forceinline__ double sqdist(const double* __restrict__ uj, const double (&y)[DP]) {
  double r2;
  if (DP == 1) {
    const double t = uj[0] - y[0];
    r2 = t * t;
  } else {
    r2 = 0.0;
#pragma unroll
    for (int k2 = 0; k2 < DP / 2; k2++) {
      const double2 u = *reinterpret_cast<const double2*>(uj + 2 * k2);
      const double t0 = u.x - y[2 * k2], t1 = u.y - y[2 * k2 + 1];
      r2 = fma(t0, t0, r2);
      r2 = fma(t1, t1, r2);
    }
    if (DP & 1) {
      const double t = uj[DP - 1] - y[DP - 1];
      r2 = fma(t, t, r2);
    }
  }
  return r2;
}

}  // namespace bode

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

}  // namespace bode

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

// ---- variance * 2^(z/64): the form the Gram matrix (K1) and the cross-covariances (K2) use -----------------------
// The coordinates are pre-scaled by kappa / (sqrt2 ell), kappa^2 = 64/ln2, so that
// exp(-r^2) = 2^(z/64) with z = -|u - y|^2 in the scaled coordinates.  k = rint(z) splits z = k + f, |f| <= 1/2
// exactly (no Cody-Waite constants needed); 2^(f/64) - 1 is a degree-5 Taylor polynomial in f (truncation 3.5e-17
// relative), the 64-entry table holds variance * 2^(j/64) and 2^(k>>6) goes into the exponent field: 9 FP64-pipe
// operations, no branches.  Valid for z in (-2^31, ~0] (the prepare kernel rejects hyper-parameters that could exceed it);
// the result saturates at variance * 2^-1000 for z < -64000.
constexpr int kExpTab64 = 64;
constexpr double kKappa = 9.608979270291599;      // sqrt(64 / ln 2)

// REP = stride between consecutive table entries (1: plain table; 16: the bank-staggered copies of layout.h, with
// `tab` already offset by the lane's copy index)
template <int REP>
__device__ __forceinline__ double rbf_exp64(double z, const double* __restrict__ tab) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double t = z + SHIFT;
  const double kf = t - SHIFT;
  const double f = z - kf;
  const int k = __double2loint(t);
  // c_i = (ln2/64)^i / i!
  double p = fma(f, 1.2417843701716925e-12, 5.732851688640402e-10);
  p = fma(p, f, 2.1173137155464776e-07);
  p = fma(p, f, 5.86490495505617e-05);
  p = fma(p, f, 0.010830424696249145);
  const double q = p * f;
  const double T = tab[(k & (kExpTab64 - 1)) * REP];
  const double res = fma(T, q, T);
  const int e = max(k >> 6, -1000);
  return __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res));
}

// N independent evaluations written stage by stage (the N dependency chains interleave in the instruction stream).
// SMEM32: `tab` is a 32-bit shared-memory address (already offset by the lane's copy): the lookup address is two
// integer operations, (k & 63) and one scaled add, instead of the four a generic pointer index compiles to.
template <int N, int REP>
__device__ __forceinline__ void rbf_exp64_n(const double* __restrict__ z, unsigned tab_smem, double* __restrict__ out) {
  const double SHIFT = 6755399441055744.0;
  double t[N], f[N], p[N], T[N];
  int k[N];
#pragma unroll
  for (int i = 0; i < N; i++) t[i] = z[i] + SHIFT;
#pragma unroll
  for (int i = 0; i < N; i++) {
    k[i] = __double2loint(t[i]);
    const unsigned addr = tab_smem + ((unsigned)(k[i] & (kExpTab64 - 1)) * (unsigned)(REP * 8));
    asm("ld.shared.f64 %0, [%1];" : "=d"(T[i]) : "r"(addr));
  }
#pragma unroll
  for (int i = 0; i < N; i++) t[i] = t[i] - SHIFT;
#pragma unroll
  for (int i = 0; i < N; i++) f[i] = z[i] - t[i];
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(f[i], 1.2417843701716925e-12, 5.732851688640402e-10);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], f[i], 2.1173137155464776e-07);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], f[i], 5.86490495505617e-05);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], f[i], 0.010830424696249145);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = p[i] * f[i];
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double res = fma(T[i], p[i], T[i]);
    const int e = max(k[i] >> 6, -1000);
    out[i] = __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res));
  }
}

}  // namespace bode

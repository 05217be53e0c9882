// K1 -- per-hyper-sample preparation (one CTA per sample).
//
// Replaces, for all S hyper-samples at once, what the reference does per (candidate, sample):
//   make_mcmc_model  bode/_sampler.py:449-462  (GPy exact inference: A = K + (noise+1e-8) I, Cholesky)
//   get_sigma_1      bode/_sampler.py:404-410  sigma_1 = bk - ek^T A^-1 ek
//   get_mu_1         bode/_sampler.py:413-422  mu_1 = (A^-1 Y)^T ek
//   xik / bk         bode/_core.py:86-101
// and additionally produces L^-1 in the tiled layout of layout.h plus w = A^-1 ek for the scoring kernel,
// and GPy's log marginal likelihood (get_likelihood, _sampler.py:221-232).
//
// Numerics: all FP64.  Unlike the reference (explicit inverse through dpotri, accurate to ~eps*cond(A)) every
// solve here goes through the triangular factor: a = L^-1 e and z = L^-1 Y come out of the factorisation as two
// bordered rows, sigma_1 = sigma_0 - a.a, mu_1 = z.a, w = L^-T a by blocked back-substitution.
//
// Algorithm: left-looking blocked Cholesky, panels of 8 columns.  Thread t owns rows i == t (mod 256) of the
// (n+2) x n bordered matrix for the whole factorisation; per panel it accumulates its row's 8 panel entries
// against the 8 staged panel rows (shared memory, broadcast reads), one warp factors the 8x8 diagonal block,
// then every row below applies the 8x8 triangular solve.  L^-1 is formed with the same panel machinery applied
// to identity rows (row r of the result = column r of L^-1), then re-tiled.
#include <cuda_runtime.h>
#include <math.h>

#include "layout.h"
#include "prepare_args.h"
#include "rbf.h"
#include "tuning.h"
#include "xik_coef.h"

namespace bode {

constexpr int kPrepThreads = 256;
constexpr int kNB = 8;


__device__ __forceinline__ double block_sum_ordered(double v, double* red) {
  // deterministic CTA-wide sum: warp shuffle tree, then warp 0 adds the 8 warp partials in order.
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kPrepThreads / 32; w++) s += red[w];
  return s;
}

// P[c] -= sum_{k in [k0,k1)} row[k] * Ljp[k][c]   (Ljp staged in smem as [k][8])
// The own-row operand comes from global memory (L2): it is fetched 16 columns at a time so that eight 16-byte loads
// are in flight per thread -- the loop is bound by that latency, not by the FMAs.
__device__ __forceinline__ void gemm_step(double (&P)[kNB], double r, const double* __restrict__ lrow) {
  const double2* lp = reinterpret_cast<const double2*>(lrow);
#pragma unroll
  for (int c2 = 0; c2 < kNB / 2; c2++) {
    const double2 l = lp[c2];
    P[2 * c2] = fma(-r, l.x, P[2 * c2]);
    P[2 * c2 + 1] = fma(-r, l.y, P[2 * c2 + 1]);
  }
}

__device__ __forceinline__ void panel_gemm_row(double (&P)[kNB], const double* __restrict__ row, const double* Ljp,
                                               int k0, int k1) {
  int k = k0;
  while ((k & 15) && k < k1) {  // peel to a 128-byte boundary of the row
    gemm_step(P, row[k], Ljp + (size_t)k * kNB);
    k++;
  }
  for (; k + 16 <= k1; k += 16) {
    double2 r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = *reinterpret_cast<const double2*>(row + k + 2 * i);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      gemm_step(P, r[i].x, Ljp + (size_t)(k + 2 * i) * kNB);
      gemm_step(P, r[i].y, Ljp + (size_t)(k + 2 * i + 1) * kNB);
    }
  }
  for (; k < k1; k++) gemm_step(P, row[k], Ljp + (size_t)k * kNB);
}

// x = P * Ld^-T for the 8x8 lower-triangular diagonal block Ld (smem, row-major [c][c']); rdiag[c] = 1 / Ld[c][c]
// (eight dependent FP64 divisions per row per panel would dominate the panel's latency; the reciprocal costs one
// extra rounding of 2^-53 per entry, well inside the factorisation's backward error)
__device__ __forceinline__ void panel_trsm_row(double (&P)[kNB], const double* Ld, const double* rdiag) {
#pragma unroll
  for (int c = 0; c < kNB; c++) {
    double t = P[c];
#pragma unroll
    for (int cc = 0; cc < c; cc++) t = fma(-P[cc], Ld[c * kNB + cc], t);
    P[c] = t * rdiag[c];
  }
}

#ifdef BODE_TIMELINE
__device__ long long g_tp[64 * 8];
#define TP(i) do { if (threadIdx.x == 0 && blockIdx.x < 64) g_tp[blockIdx.x * 8 + (i)] = clock64(); } while (0)
cudaError_t debug_copy_prepare_timeline(long long* host) { return cudaMemcpyFromSymbol(host, g_tp, sizeof(g_tp)); }
#else
#define TP(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kPrepThreads, 1) k_prepare(PrepArgs a) {
  extern __shared__ __align__(16) double sm[];
  if ((int)blockIdx.x >= a.S) {  // Chebyshev series of the xik factors (blocks beyond the S factorisation CTAs)
    const int per = (a.d + 3) / 4, b = (int)blockIdx.x - a.S;
    xik_coef_block(a.hyp, a.ndim, a.d, a.ws, a.L, b / per, 4 * (b % per), sm, sm + 4 * kXcNodes);
    return;
  }
  const int s = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, d = a.d, npad = a.L.npad;
  TP(0);
  double* Ljp = sm;                        // [npad][8] staged panel rows
  double* Ld = Ljp + (size_t)npad * kNB;   // [8][8] diagonal block
  double* rdiag = Ld + kNB * kNB;          // [8] reciprocals of the diagonal of Ld
  double* red = rdiag + kNB;               // [8] reduction scratch
  double* ck = red + 8;                    // [16]
  double* cku = ck + kCkPad;               // [16] kappa * ck: scale of U (rbf.h)
  double* tab = cku + kCkPad;              // [64] variance * 2^(j/64)
  double* vec = tab + kTab64;              // [npad] back-substitution right-hand side
  __shared__ int fail;

  double* aux = reinterpret_cast<double*>(a.ws + a.L.off_aux) + (size_t)s * a.L.aux_doubles;
  double* Mt = reinterpret_cast<double*>(a.ws + a.L.off_m) + (size_t)s * a.L.m_doubles;
  double* Lw = reinterpret_cast<double*>(a.ws + a.L.off_lw) + (size_t)s * a.L.lw_doubles;
  double* Xw = reinterpret_cast<double*>(a.ws + a.L.off_xw) + (size_t)s * a.L.xw_doubles;
  double* Ut = aux + aux_ut(a.L);   // U in A-fragment tiles (layout.h)
  const int dk = a.L.dk;

  const double* h = a.hyp + (size_t)s * a.ndim;
  const double ss = h[0];
  const double noise = (a.ndim == d + 2) ? h[d + 1] * h[d + 1] : a.nugget_var;
  const double SQRT2 = 1.4142135623730951, SQRTPI = 1.7724538509055159;
  if (tid < kCkPad) {
    const double c = (tid < d) ? 1.0 / (SQRT2 * h[1 + tid]) : 0.0;
    ck[tid] = c;
    cku[tid] = kKappa * c;
  }
  if (tid >= 32 && tid < 32 + kTab64) tab[tid - 32] = ss * exp2((double)(tid - 32) / 64.0);
  if (tid == 0) {
    // hyper-parameter sanity: the scoring kernel's exponent arithmetic assumes a sane variance scale
    int bad = !(ss > 9.5e-7 && ss < 1.0e6) || !(noise >= 0.0 && noise < 1.0e6);
    // ... and that the scaled squared distances stay below 2^31 (rbf.h): sum_k kappa^2 / (2 ell_k^2) < 2e9, i.e.
    // lengthscales of at least ~2e-4 of the unit box
    double c2 = 0.0;
    for (int k = 0; k < d; k++) {
      bad |= !(h[1 + k] > 1e-300 && h[1 + k] < 1e300);
      c2 += (kKappa * kKappa) / (2.0 * h[1 + k] * h[1 + k]);
    }
    bad |= !(c2 < 2.0e9);
    fail = bad ? -1 : 0;
  }
  __syncthreads();
  if (tid < kCkPad) {
    aux[aux_ck() + tid] = ck[tid];
    aux[aux_ell() + tid] = (tid < d) ? h[1 + tid] : 1.0;
    aux[aux_cks() + tid] = 2.0 * cku[tid];
  }
  for (int idx = tid; idx < kTab64 * kTabRep; idx += kPrepThreads) aux[aux_tabr() + idx] = tab[idx / kTabRep];

  // ---- prescaled, centred coordinates U (fragment-tiled), -|U_j|^2, bordered rows e = xik(X) (row n) and Y (row n+1)
  for (int idx = tid; idx < npad * 4 * dk; idx += kPrepThreads) {
    const int blk = idx >> 5, ln = idx & 31;          // block (R, i), lane = rr*4 + kk
    const int R = blk / dk, i = blk - R * dk;
    const int j = 8 * R + (ln >> 2), k = 4 * i + (ln & 3);
    Ut[idx] = (j < n && k < d) ? (a.X[(size_t)j * d + k] - 0.5) * cku[k] : 0.0;
  }
  for (int j = tid; j < npad; j += kPrepThreads) {  // fixed order over k
    double nun = 0.0;
    if (j < n)
      for (int k = 0; k < d; k++) {
        const double u = (a.X[(size_t)j * d + k] - 0.5) * cku[k];
        nun = fma(-u, u, nun);
      }
    aux[aux_nw() + 2 * j] = nun;
  }
  for (int j = tid; j < npad; j += kPrepThreads) {
    double e = 0.0, y = 0.0;
    if (j < n) {
      // _core.py:86-92: ss * prod_k sqrt(pi)/sqrt(2) * ell_k * (erf((1-x)/(sqrt2 ell)) - erf((0-x)/(sqrt2 ell)))
      double p = 1.0;
      for (int k = 0; k < d; k++) {
        const double x = a.X[(size_t)j * d + k];
        p *= (SQRTPI / SQRT2) * h[1 + k] * (erf((1.0 - x) * ck[k]) - erf((0.0 - x) * ck[k]));
      }
      e = a.e_ovr ? a.e_ovr[(size_t)s * n + j] : ss * p;
      y = a.Y[j];
    }
    Lw[(size_t)n * npad + j] = e;
    Lw[(size_t)(n + 1) * npad + j] = y;
  }
  __syncthreads();
  // ---- A = K + (noise + jitter) I, lower triangle, rows 0..n-1 ------------------------------------------------
  // strictly-lower entries, flat triangular index so that all threads carry the same number of exp() calls
  // (the observations are staged in the panel buffer, free until the factorisation starts, when they fit)
  const double jit = a.jitter_s ? a.jitter_s[s] : a.jitter;
  const bool x_smem = d <= kNB;
  if (x_smem) {
    for (int idx = tid; idx < n * d; idx += kPrepThreads) Ljp[idx] = a.X[idx];
    __syncthreads();
  }
  const double* Xs = x_smem ? Ljp : a.X;
  const int ntri = n * (n - 1) / 2;
  for (int idx = tid; idx < ntri; idx += kPrepThreads) {
    int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)idx)) * 0.5);
    while (i * (i - 1) / 2 > idx) i--;
    while ((i + 1) * i / 2 <= idx) i++;
    const int j = idx - i * (i - 1) / 2;
    const double* xi = Xs + (size_t)i * d;
    const double* xj = Xs + (size_t)j * d;
    double z = 0.0;
    for (int k = 0; k < d; k++) {
      const double t = (xi[k] - xj[k]) * cku[k];
      z = fma(-t, t, z);
    }
    Lw[(size_t)i * npad + j] = rbf_exp64<1>(z, tab);
  }
  for (int idx = tid; idx < n * npad; idx += kPrepThreads) {
    const int i = idx / npad, j = idx - i * npad;
    if (j == i) Lw[idx] = ss + (noise + jit);  // GPy: diag.add(Ky, variance + 1e-8)
    else if (j > i) Lw[idx] = 0.0;
  }
  __syncthreads();

  TP(1);
  // ---- phase 1: bordered left-looking Cholesky -----------------------------------------------------------------
  const int nrows = n + 2;
  int info = fail;  // -1: rejected hyper-parameters
  for (int j0 = 0; j0 < n && info == 0; j0 += kNB) {
    const int nb = min(kNB, n - j0);
    // stage panel rows j0..j0+7, columns [0, j0), as [k][8]
    for (int idx = tid; idx < j0 * kNB; idx += kPrepThreads) {
      const int c = idx / j0, k = idx - c * j0;  // consecutive threads read consecutive k of one row
      Ljp[(size_t)k * kNB + c] = (c < nb) ? Lw[(size_t)(j0 + c) * npad + k] : 0.0;
    }
    __syncthreads();
    // each owned row >= j0: P = A[i][j0..j0+8) - L[i][:j0] . L[j0+c][:j0]
    for (int i = tid; i < nrows; i += kPrepThreads) {
      if (i < j0) continue;
      double* row = Lw + (size_t)i * npad;
      double P[kNB];
#pragma unroll
      for (int c = 0; c < kNB; c++) P[c] = (c < nb) ? row[j0 + c] : 0.0;
      panel_gemm_row(P, row, Ljp, 0, j0);
#pragma unroll
      for (int c = 0; c < kNB; c++)
        if (c < nb && (i >= j0 + nb || c <= i - j0)) row[j0 + c] = P[c];  // keep the strict upper part zero
    }
    __syncthreads();
    // 8x8 diagonal block: lane c of warp 0 owns block row c
    if (tid < 32) {
      const int c = tid & 7;
      double B[kNB];
#pragma unroll
      for (int cc = 0; cc < kNB; cc++) B[cc] = (c < nb && cc <= c && cc < nb) ? Lw[(size_t)(j0 + c) * npad + j0 + cc] : ((cc == c) ? 1.0 : 0.0);
      int bad = 0;
#pragma unroll
      for (int cc = 0; cc < kNB; cc++) {
        // pivot of column cc lives in lane cc
        double piv = __shfl_sync(0xffffffffu, B[cc], cc);
        if (!(piv > 0.0) && bad == 0) bad = j0 + cc + 1;
        const double dd = sqrt(piv);
        if (c == cc) B[cc] = dd;
        else if (c > cc) B[cc] = B[cc] / dd;
        // trailing update of this lane's row: B[c'] -= l_c,cc * l_c',cc for cc < c' <= c
#pragma unroll
        for (int c2 = cc + 1; c2 < kNB; c2++) {
          const double lc2 = __shfl_sync(0xffffffffu, B[cc], c2);  // l_{c2,cc} (valid once lane c2 divided)
          if (c2 <= c) B[c2] = fma(-B[cc], lc2, B[c2]);
        }
      }
      if (tid < 8) {
#pragma unroll
        for (int cc = 0; cc < kNB; cc++) {
          Ld[c * kNB + cc] = (cc <= c) ? B[cc] : 0.0;
          if (cc == c) rdiag[c] = 1.0 / B[cc];
          if (c < nb && cc <= c && cc < nb) Lw[(size_t)(j0 + c) * npad + j0 + cc] = B[cc];
        }
      }
      if (tid == 0 && bad != 0 && bad <= n) fail = bad;
    }
    __syncthreads();
    if (fail) { info = fail; break; }
    // rows below the diagonal block (incl. the two bordered rows): triangular solve against Ld
    for (int i = tid; i < nrows; i += kPrepThreads) {
      if (i < j0 + nb) continue;
      double* row = Lw + (size_t)i * npad;
      double P[kNB];
#pragma unroll
      for (int c = 0; c < kNB; c++) P[c] = (c < nb) ? row[j0 + c] : 0.0;
      panel_trsm_row(P, Ld, rdiag);
#pragma unroll
      for (int c = 0; c < kNB; c++)
        if (c < nb) row[j0 + c] = P[c];
    }
    __syncthreads();
  }

  TP(2);
  double* cst = aux;  // consts
  if (info) {
    // failed factorisation: poison the sample (NaN propagates to every EKLD of this sample), zero the factors.
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (tid == 0) {
      cst[C_SS] = ss; cst[C_NOISE] = noise; cst[C_SIGMA0] = qnan; cst[C_SIGMA1] = qnan; cst[C_MU1] = qnan;
      cst[C_LOGDET] = qnan; cst[C_QUAD] = qnan; cst[C_LOGLIK] = -INFINITY;
      a.info[s] = info;
    }
    for (int j = tid; j < npad; j += kPrepThreads) aux[aux_nw() + 2 * j + 1] = 0.0;
    for (size_t idx = tid; idx < a.L.m_doubles; idx += kPrepThreads) Mt[idx] = 0.0;
    return;
  }

  // ---- scalars: sigma0 (bk), sigma1, mu1, logdet, quad ---------------------------------------------------------
  const double* arow = Lw + (size_t)n * npad;
  const double* zrow = Lw + (size_t)(n + 1) * npad;
  double paa = 0.0, pza = 0.0, pzz = 0.0, pld = 0.0;
  for (int j = tid; j < n; j += kPrepThreads) {
    const double av = arow[j], zv = zrow[j];
    paa = fma(av, av, paa);
    pza = fma(zv, av, pza);
    pzz = fma(zv, zv, pzz);
    pld += log(Lw[(size_t)j * npad + j]);
  }
  const double saa = block_sum_ordered(paa, red);
  const double sza = block_sum_ordered(pza, red);
  const double szz = block_sum_ordered(pzz, red);
  const double sld = 2.0 * block_sum_ordered(pld, red);
  if (tid == 0) {
    // _core.py:95-101
    double p = 1.0;
    for (int k = 0; k < d; k++) {
      const double ell = h[1 + k], zk = 1.0 / (SQRT2 * ell);
      p *= 2.0 * SQRTPI * (ell * ell) * (zk * erf(zk) + exp(-(zk * zk)) / SQRTPI - 1.0 / SQRTPI);
    }
    const double sigma0 = a.sigma0_ovr ? a.sigma0_ovr[s] : ss * p;
    cst[C_SS] = ss; cst[C_NOISE] = noise; cst[C_SIGMA0] = sigma0; cst[C_SIGMA1] = sigma0 - saa; cst[C_MU1] = sza;
    cst[C_LOGDET] = sld; cst[C_QUAD] = szz;
    cst[C_LOGLIK] = 0.5 * (-(double)n * 1.8378770664093453 - sld - szz);  // log(2 pi)
    a.info[s] = 0;
  }

  if (a.loglik_only) return;  // the MCMC likelihood path needs only consts[C_LOGLIK]

  TP(3);
  // ---- w = L^-T a : blocked back-substitution, 8 unknowns per step ---------------------------------------------
  for (int j = tid; j < npad; j += kPrepThreads) vec[j] = (j < n) ? arow[j] : 0.0;
  __syncthreads();
  for (int j0 = ((n - 1) / kNB) * kNB; j0 >= 0; j0 -= kNB) {
    const int nb = min(kNB, n - j0);
    if (tid < kNB * kNB) {
      const int c = tid / kNB, cc = tid % kNB;
      Ld[tid] = (c < nb && cc <= c) ? Lw[(size_t)(j0 + c) * npad + j0 + cc] : ((c == cc) ? 1.0 : 0.0);
    }
    __syncthreads();
    if (tid == 0) {
      // solve Ld^T w_blk = vec_blk sequentially (8x8), Ld = L[j0.., j0..] staged in smem
      double wb[kNB];
      for (int c = nb - 1; c >= 0; c--) {
        double t = vec[j0 + c];
        for (int c2 = c + 1; c2 < nb; c2++) t = fma(-Ld[c2 * kNB + c], wb[c2], t);
        wb[c] = t / Ld[c * kNB + c];
      }
      for (int c = 0; c < nb; c++) vec[j0 + c] = wb[c];
    }
    __syncthreads();
    // vec[k] -= sum_c w[j0+c] * L[j0+c][k] for k < j0
    for (int k = tid; k < j0; k += kPrepThreads) {
      double t = vec[k];
      for (int c = 0; c < nb; c++) t = fma(-vec[j0 + c], Lw[(size_t)(j0 + c) * npad + k], t);
      vec[k] = t;
    }
    __syncthreads();
  }
  for (int j = tid; j < npad; j += kPrepThreads) aux[aux_nw() + 2 * j + 1] = (j < n) ? vec[j] : 0.0;

  TP(4);
  // ---- phase 2: X = I L^-T  (row r of X = column r of L^-1), same panel machinery -------------------------------
  for (int j0 = 0; j0 < n; j0 += kNB) {
    const int nb = min(kNB, n - j0);
    __syncthreads();
    for (int idx = tid; idx < j0 * kNB; idx += kPrepThreads) {
      const int c = idx / j0, k = idx - c * j0;
      Ljp[(size_t)k * kNB + c] = (c < nb) ? Lw[(size_t)(j0 + c) * npad + k] : 0.0;
    }
    if (tid < kNB * kNB) {
      const int c = tid / kNB, cc = tid % kNB;
      const double v = (c < nb && cc <= c) ? Lw[(size_t)(j0 + c) * npad + j0 + cc] : ((c == cc) ? 1.0 : 0.0);
      Ld[tid] = v;
      if (c == cc) rdiag[c] = 1.0 / v;
    }
    __syncthreads();
    for (int r = tid; r < j0 + nb; r += kPrepThreads) {
      double* row = Xw + (size_t)r * npad;
      double P[kNB];
#pragma unroll
      for (int c = 0; c < kNB; c++) P[c] = (r == j0 + c) ? 1.0 : 0.0;
      if (r < j0) panel_gemm_row(P, row, Ljp, r, j0);
      panel_trsm_row(P, Ld, rdiag);
#pragma unroll
      for (int c = 0; c < kNB; c++)
        if (c < nb) row[j0 + c] = P[c];
    }
  }
  __syncthreads();
  TP(5);
  // ---- re-tile: block (R, g) element (rr, kk) = M[8R+rr][4g+kk] = X[4g+kk][8R+rr] --------------------------------
  const int nrt = a.L.nrt, ng = a.L.ng;
  for (int g = 0; g < ng; g++) {
    const size_t goff = m_goff(nrt, g);
    const int r0 = g >> 1;
    const int cnt = (nrt - r0) * 32;
    for (int idx = tid; idx < cnt; idx += kPrepThreads) {
      const int blk = idx >> 5, e = idx & 31;
      const int kk = e >> 3, rr = e & 7;  // consecutive threads walk rr (contiguous in X rows)
      const int i = 8 * (r0 + blk) + rr, j = 4 * g + kk;
      double v = 0.0;
      if (i < n && j < n && j <= i) v = Xw[(size_t)j * npad + i];
      Mt[goff + (size_t)blk * 32 + rr * 4 + kk] = v;
    }
  }
  TP(6);
}


size_t prepare_smem_bytes(const Layout& L) {
  const size_t fact = (size_t)L.npad * kNB + kNB * kNB + kNB + 8 + 2 * kCkPad + kTab64 + L.npad;
  const size_t coef = 8 * (size_t)kXcNodes;  // the series blocks of the same launch
  return sizeof(double) * (fact > coef ? fact : coef);
}

size_t prepare_tiles_smem_bytes(const Layout& L);                          // prepare_tiles.cu
size_t smem_optin_bytes();                                                 // prepare_tiles.cu
cudaError_t launch_prepare_tiles(const PrepArgs& a, cudaStream_t st);      // prepare_tiles.cu

cudaError_t launch_prepare(const PrepArgs& a, cudaStream_t st) {
  // n <= 224: the bordered triangle fits in shared memory and the factorisation runs on the FP64 tensor pipe
  // (prepare_tiles.cu); larger n (or BODE_PREP_VARIANT=0): the global-memory kernel below
  if (tuning().prep_variant != 0 && a.L.nrt <= 28 && prepare_tiles_smem_bytes(a.L) + 64 <= smem_optin_bytes())
    return launch_prepare_tiles(a, st);
  const size_t smem = prepare_smem_bytes(a.L);
  static thread_local int attr_dev = -1;
  static thread_local size_t attr_smem = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_dev || smem > attr_smem) {
    e = cudaFuncSetAttribute(k_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
    attr_smem = smem;
  }
  // S factorisation CTAs, plus (scoring state only) S * ceil(d/4) small CTAs for the xik series
  const int extra = a.loglik_only ? 0 : a.S * ((a.d + 3) / 4);
  k_prepare<<<a.S + extra, kPrepThreads, smem, st>>>(a);
  return cudaGetLastError();
}

// ---- tiny helpers on the prepared state ------------------------------------------------------------------------
__global__ void k_sample_terms(const char* ws, Layout L, double* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L.S) return;
  const double* aux = reinterpret_cast<const double*>(ws + L.off_aux) + (size_t)s * L.aux_doubles;
  for (int i = 0; i < kConsts; i++) out[(size_t)s * kConsts + i] = aux[i];
}

__global__ void k_mu_sigma(const char* ws, Layout L, double* out) {
  // get_mu_sigma (_sampler.py:434-444): sequential accumulation over hyper-samples, then / S
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double mu = 0.0, sg = 0.0;
  for (int s = 0; s < L.S; s++) {
    const double* aux = reinterpret_cast<const double*>(ws + L.off_aux) + (size_t)s * L.aux_doubles;
    mu += aux[C_MU1];
    sg += aux[C_SIGMA1];
  }
  out[0] = mu / (double)L.S;
  out[1] = sg / (double)L.S;
}

__global__ void k_loglik_gather(const char* ws, Layout L, double* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L.S) return;
  out[s] = (reinterpret_cast<const double*>(ws + L.off_aux) + (size_t)s * L.aux_doubles)[C_LOGLIK];
}

cudaError_t launch_loglik_gather(const void* ws, const Layout& L, double* out, cudaStream_t st) {
  k_loglik_gather<<<(L.S + 127) / 128, 128, 0, st>>>(static_cast<const char*>(ws), L, out);
  return cudaGetLastError();
}

cudaError_t launch_sample_terms(const void* ws, const Layout& L, double* out, cudaStream_t st) {
  k_sample_terms<<<(L.S + 127) / 128, 128, 0, st>>>(static_cast<const char*>(ws), L, out);
  return cudaGetLastError();
}
cudaError_t launch_mu_sigma(const void* ws, const Layout& L, double* out, cudaStream_t st) {
  k_mu_sigma<<<1, 32, 0, st>>>(static_cast<const char*>(ws), L, out);
  return cudaGetLastError();
}

}  // namespace bode

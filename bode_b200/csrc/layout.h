// Device-workspace layout shared by the prepare (K1) and scoring (K2-K4) kernels and the host C ABI.
//
// Per hyper-sample s the workspace holds
//   aux[s]   : one contiguous block (copied to shared memory with a single bulk copy by the scoring kernel)
//                consts[8]  = {variance, noise_var, sigma0, sigma1, mu1, logdet, Y^T A^-1 Y, loglik}
//                ck[16]     = 1/(sqrt(2) ell_k)            (zero padded; arguments of erf in xik)
//                ell[16]    = ell_k                        (one padded)
//                cks[16]    = 2 kappa ck                   (zero padded; kappa^2 = 64/ln 2, see rbf.h)
//                tabr[64][16] = variance * 2^(j/64), each entry replicated 16 times so that lane l reads copy l & 15:
//                               the per-lane table lookups of a warp never collide on a shared-memory bank
//                nw[npad][2] = {-|U_j|^2, (A^-1 e)_j}      (zero padded)
//                Ut[nrt][dk][32] = U = kappa ck[k] * (X[j][k] - 1/2) in mma.m8n8k4 A-fragment order:
//                               entry lane = rr*4 + kk of block (R, i) is U[8R + rr][4i + kk]   (zero padded)
//              so that the cross-covariance tile of 8 observations x 8 candidates is
//                  z = -|U_j|^2 - |y|^2 + 2 U_j . y   (dk DMMAs, y = the candidate scaled like U),   k = tab-exp2(z / 64).
//   M[s]     : L_s^-1 in the "k4-group major, live row-tile suffix" tiling the DMMA mainloop consumes:
//                for g = 0..ng-1 (columns j = 4g..4g+3), for R = g/2..nrt-1 (rows 8R..8R+7): a 256-byte block
//                holding M[8R+rr][4g+kk] at [rr*4 + kk] -- exactly the A-fragment order of mma.m8n8k4.f64.
//                Row tiles R < g/2 lie strictly above the diagonal and are not stored.
//   Lw[s]    : (n+2) x npad row-major scratch: Cholesky factor rows 0..n-1, then a = L^-1 e and z = L^-1 Y.
//   Xw[s]    : n x npad row-major scratch: M^T (row r = column r of L^-1).
//   xc[s]    : per dimension k a block of kXcStride doubles: the even Chebyshev series of the xik factor
//                F_k(x) = sqrt(pi/2) ell_k (erf((1-x) c_k) + erf(x c_k)) = a_0/2 + sum_j a_j T_j(w),  w = 2 (2x-1)^2 - 1
//                (F_k is symmetric about x = 1/2): [n_terms, a_0/2, a_1, ..., a_(kXcMax-1)]; n_terms = 0 means the
//                series did not converge within kXcMax terms (very short lengthscale) and the factor is evaluated
//                with erf directly.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BODE_HD __host__ __device__ __forceinline__
#else
#define BODE_HD inline
#endif

namespace bode {

constexpr int kConsts = 8;
constexpr int kCkPad = 16;
constexpr int kTabPad = 32;   // table of rbf_exp (quadrature row means)
constexpr int kTab64 = 64;    // table of rbf_exp64 (Gram matrix and cross-covariances)
constexpr int kXcMax = 48;     // longest even Chebyshev series kept for an xik factor
constexpr int kXcNodes = 64;   // Gauss-Chebyshev nodes the series is computed from
constexpr int kXcStride = 50;  // doubles per (sample, dimension) block of the xc section
constexpr int kTabRep = 16;   // bank-staggered copies of each table entry in the aux block
enum ConstIdx { C_SS = 0, C_NOISE = 1, C_SIGMA0 = 2, C_SIGMA1 = 3, C_MU1 = 4, C_LOGDET = 5, C_QUAD = 6, C_LOGLIK = 7 };

struct Layout {
  int n, d, S;
  int nrt;   // row tiles of 8
  int npad;  // 8 * nrt
  int ng;    // k4 groups = 2 * nrt
  int dk;    // k4 steps of the cross-covariance dot product = ceil(d / 4)
  size_t aux_doubles;  // per sample
  size_t m_doubles;    // per sample
  size_t lw_doubles;   // per sample
  size_t xw_doubles;   // per sample
  size_t xc_doubles;   // per sample
  size_t off_aux, off_m, off_lw, off_xw, off_xc, total_bytes;  // byte offsets of the five sections
};

BODE_HD size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

BODE_HD Layout make_layout(int n, int d, int S) {
  Layout L;
  L.n = n; L.d = d; L.S = S;
  L.nrt = (n + 7) / 8;
  L.npad = 8 * L.nrt;
  L.ng = 2 * L.nrt;
  L.dk = (d + 3) / 4;
  L.aux_doubles = (size_t)kConsts + 3 * kCkPad + kTab64 * kTabRep + 2 * (size_t)L.npad + (size_t)L.npad * 4 * L.dk;
  L.m_doubles = (size_t)32 * L.nrt * (L.nrt + 1);
  L.lw_doubles = (size_t)(n + 2) * L.npad;
  L.xw_doubles = (size_t)n * L.npad;
  L.off_aux = 0;
  L.off_m = align_up(L.off_aux + sizeof(double) * L.aux_doubles * S, 256);
  L.off_lw = align_up(L.off_m + sizeof(double) * L.m_doubles * S, 256);
  L.off_xw = align_up(L.off_lw + sizeof(double) * L.lw_doubles * S, 256);
  L.xc_doubles = (size_t)kCkPad * kXcStride;
  L.off_xc = align_up(L.off_xw + sizeof(double) * L.xw_doubles * S, 256);
  L.total_bytes = align_up(L.off_xc + sizeof(double) * L.xc_doubles * S, 256);
  return L;
}

// aux sub-offsets (doubles)
BODE_HD int aux_ck() { return kConsts; }
BODE_HD int aux_ell() { return kConsts + kCkPad; }
BODE_HD int aux_cks() { return kConsts + 2 * kCkPad; }
BODE_HD int aux_tabr() { return kConsts + 3 * kCkPad; }
BODE_HD int aux_nw() { return kConsts + 3 * kCkPad + kTab64 * kTabRep; }   // nw and Ut may stay in global memory (large n)
BODE_HD int aux_ut(const Layout& L) { return aux_nw() + 2 * L.npad; }

// offset (doubles) of k4-group g inside a sample's tiled M; block (R, g), R >= g/2, sits at m_goff(g) + (R - g/2)*32
BODE_HD size_t m_goff(int nrt, int g) {
  const size_t a = (size_t)(g >> 1), b = (size_t)(g & 1);
  return 32 * (2 * (a * nrt - a * (a - 1) / 2) + b * (nrt - a));
}

}  // namespace bode

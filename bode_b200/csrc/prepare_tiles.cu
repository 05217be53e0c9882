// K1, shared-memory tile variant -- per-hyper-sample preparation for n <= 224 observations (one CTA per sample).
//
// Same contract as k_prepare (prepare.cu; reference make_mcmc_model bode/_sampler.py:449-462, get_sigma_1 :404-410,
// get_mu_1 :413-422, xik / bk _core.py:86-101, GPy log marginal likelihood :221-232), different machine mapping:
// the whole bordered lower triangle lives in shared memory as 8 x 8 tiles and every O(n^3) step runs on the FP64
// tensor pipe (mma.sync.m8n8k4.f64).  k_prepare keeps the factor in global memory and walks it thread-per-row; at
// config 3 (n = 200) that is 0.56 ms of L2 latency and dependent FMA chains on 64 of 148 SMs -- replicated on every rank,
// it is what held strong scaling at 65 % on 8 GPUs.
//
// Tile storage ("fragment order"): tile (I, J), I >= J, is 64 doubles, two 8 x 4 halves; element [r][c] sits at
//     (c >> 2) * 32 + r * 4 + (c & 3).
// Lane l reading word h * 32 + l gets element [l >> 2][(l & 3) + 4h]: that is at once the A fragment of the tile's
// columns 4h..4h+3 and the B fragment of the TRANSPOSED tile's rows 4h..4h+3 (conflict-free 8-byte loads); the
// accumulator pair of lane (r = l >> 2, q = l & 3), elements [r][2q], [r][2q+1], is one aligned 16-byte word; and the
// two halves are byte-for-byte the blocks (R = I, g = 2J), (R = I, g = 2J + 1) of the tiled L^-1 the scoring kernel
// streams (layout.h), so the result leaves shared memory as plain 256-byte copies.
// Row tile T = nrt carries the two bordered rows e = xik(X) (row 0) and Y (row 1): a = L^-1 e and z = L^-1 Y fall out
// of the factorisation like any other row.
//
// Schedule (8 warps):
//   Gram      tiles round-robin over the warps: GEMM-form squared distance on the tensor pipe against the fragment-tiled
//             scaled observations (the arithmetic of the scoring kernel's cross-covariance) + the shared table exp2.
//   Cholesky  left-looking over block columns J: every warp accumulates  sum_K L[I][K] L[J][K]^T  for its tiles of the
//             panel (I - J == 1 + (warp - 1) mod 7; up to 4 tiles per warp, independent DMMA chains), warp 0 owns the
//             diagonal tile: 8 x 8 Cholesky with lane c holding row c (rsqrt + one Newton step per pivot, no division),
//             then W_J = L_JJ^-1 by forward substitution, stored IN PLACE of L_JJ (the diagonal tile is never an
//             update operand; its diagonal goes to ldiag for the log-determinant).  After a barrier the other warps
//             finish  L[I][J] = C W_J^T  (accumulator -> A fragment by four register shuffles).  Two barriers per panel.
//   Inverse   M = L^-1 row by row, in place:  N[I][K] = W_I L[I][K]  (K < I), barrier,
//             M[I][J] = - sum_{K=J}^{I-1} N[I][K] M[K][J]  in registers, barrier, store.  Two barriers per row.
//   w = M^T a (thread per column), scalars, then the tiles are copied out in the scoring kernel's layout.
#include <cuda_runtime.h>
#include <math.h>

#include "layout.h"
#include "prepare_args.h"
#include "rbf.h"
#include "xik_coef.h"

namespace bode {

constexpr int kPtThreads = 256;
constexpr int kPtWarps = kPtThreads / 32;
constexpr int kPtHeld = 4;  // tiles a warp holds in registers at once (nrt <= 28: ceil(28 / 7) = 4)

#ifdef BODE_TIMELINE
__device__ long long g_tpt[64 * 16];
#define TPT(i) do { if (threadIdx.x == 0 && blockIdx.x < 64) g_tpt[blockIdx.x * 16 + (i)] = clock64(); } while (0)
#define TPT_ACC_BEGIN() long long tacc0_ = clock64()
#define TPT_ACC(i, w) do { tacc_[(i) - 8] += clock64() - tacc0_; } while (0)
#define TPT_ACC_FLUSH() do { if (lane == 0 && blockIdx.x < 64) { if (warp == 0) { g_tpt[blockIdx.x * 16 + 9] = tacc_[1]; g_tpt[blockIdx.x * 16 + 13] = tacc_[5]; } \
                                                            if (warp == 1) { g_tpt[blockIdx.x * 16 + 8] = tacc_[0]; g_tpt[blockIdx.x * 16 + 14] = tacc_[6]; } } } while (0)
cudaError_t debug_copy_prepare_tiles_timeline(long long* host) { return cudaMemcpyFromSymbol(host, g_tpt, sizeof(g_tpt)); }
#else
#define TPT(i) do { } while (0)
#define TPT_ACC_BEGIN() do { } while (0)
#define TPT_ACC(i, w) do { } while (0)
#define TPT_ACC_FLUSH() do { } while (0)
#endif

__device__ __forceinline__ void pt_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int pt_pos(int r, int c) { return (c >> 2) * 32 + r * 4 + (c & 3); }
__device__ __forceinline__ int pt_slot(int I, int J) { return (I * (I + 1) / 2 + J) * 64; }

__device__ __forceinline__ double pt_block_sum(double v, double* red) {
  // deterministic CTA-wide sum: warp shuffle tree, then the 8 warp partials in order (same scheme as prepare.cu)
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kPtWarps; w++) s += red[w];
  return s;
}

// ---- statically specialised inner loops: a predicated-off DMMA still occupies the tensor pipe, so the number of live
// tiles of a warp is a template parameter, never a predicate ------------------------------------------------------------
// sum_K L[I_m][K] L[J][K]^T for CNT tiles of block column J; the two k4 halves go to separate accumulators
template <int CNT>
__device__ __forceinline__ void pt_update(double (&c0)[kPtHeld], double (&c1)[kPtHeld], const double* __restrict__ rowJ,
                                          const double* const (&rowm)[kPtHeld], int J) {
  // (splitting further by the parity of K was measured and does not pay: the loop is not DMMA-latency bound)
  double d0[CNT], d1[CNT];
#pragma unroll
  for (int m = 0; m < CNT; m++) { d0[m] = 0.0; d1[m] = 0.0; }
  for (int K = 0; K < J; K++) {
    const double b0 = rowJ[K * 64], b1 = rowJ[K * 64 + 32];
    double a0[CNT], a1[CNT];
#pragma unroll
    for (int m = 0; m < CNT; m++) {
      a0[m] = rowm[m][K * 64];
      a1[m] = rowm[m][K * 64 + 32];
    }
#pragma unroll
    for (int m = 0; m < CNT; m++) pt_dmma(c0[m], c1[m], a0[m], b0);
#pragma unroll
    for (int m = 0; m < CNT; m++) pt_dmma(d0[m], d1[m], a1[m], b1);
  }
#pragma unroll
  for (int m = 0; m < CNT; m++) { c0[m] += d0[m]; c1[m] += d1[m]; }
}

// rows K in [Kbeg, Kend) of  sum_K N[I][K] M[K][J_m]  for the first CNT columns J_m = warp + 8m of this warp
template <int CNT>
__device__ __forceinline__ void pt_mpass(double (&m0)[kPtHeld], double (&m1)[kPtHeld], double (&e0)[kPtHeld], double (&e1)[kPtHeld],
                                         const double* __restrict__ rowI, const double* __restrict__ tiles, int boff, int warp,
                                         int Kbeg, int Kend) {
  for (int K = Kbeg; K < Kend; K++) {
    const double a0 = rowI[K * 64], a1 = rowI[K * 64 + 32];
    const double* rowK = tiles + pt_slot(K, warp) + boff;
    double b0[CNT], b1[CNT];
#pragma unroll
    for (int m = 0; m < CNT; m++) {
      b0[m] = rowK[m * (kPtWarps * 64)];
      b1[m] = rowK[m * (kPtWarps * 64) + 16];
    }
#pragma unroll
    for (int m = 0; m < CNT; m++) pt_dmma(m0[m], m1[m], a0, b0[m]);
#pragma unroll
    for (int m = 0; m < CNT; m++) pt_dmma(e0[m], e1[m], a1, b1[m]);
  }
}

__global__ void __launch_bounds__(kPtThreads, 1) k_prepare_tiles(PrepArgs a) {
  extern __shared__ __align__(16) double sm[];
  if ((int)blockIdx.x >= a.S) {  // Chebyshev series of the xik factors (blocks beyond the S factorisation CTAs)
    const int per = (a.d + 3) / 4, b = (int)blockIdx.x - a.S;
    xik_coef_block(a.hyp, a.ndim, a.d, a.ws, a.L, b / per, 4 * (b % per), sm, sm + 4 * 64);
    return;
  }
  const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.n, d = a.d, npad = a.L.npad, T = a.L.nrt;
#ifdef BODE_TIMELINE
  if (tid < 16 && blockIdx.x < 64) g_tpt[blockIdx.x * 16 + tid] = 0;
  __syncthreads();
#endif
  TPT(0);
  const int r = lane >> 2, q = lane & 3;
  const int ntiles = T * (T + 1) / 2 + T;  // lower triangle + the bordered row tile
  double* tiles = sm;
  double* ldiag = tiles + (size_t)ntiles * 64;  // [npad] diagonal of L
  double* av = ldiag + npad;                    // [npad] a = L^-1 e
  double* zv = av + npad;                       // [npad] z = L^-1 Y
  double* wv = zv + npad;                       // [npad] w = L^-T a
  double* ck = wv + npad;                       // [16]
  double* cku = ck + kCkPad;                    // [16]
  double* tab = cku + kCkPad;                   // [64] variance * 2^(j/64)
  double* red = tab + kTab64;                   // [8]
  double* nunv = red + 8;                       // [npad] -|U_j|^2
  unsigned short* tij = reinterpret_cast<unsigned short*>(nunv + npad);  // [T (T + 1) / 2] (I << 8) | J of lower tile t
  const int nlow = T * (T + 1) / 2;
  double* Us = nunv + npad + ((nlow + 3) >> 2);   // [nrt][dk][32] scaled observations in fragment tiles, staged when they fit
  const bool u_smem = a.x_smem != 0;
  __shared__ int fail;

  double* aux = reinterpret_cast<double*>(a.ws + a.L.off_aux) + (size_t)s * a.L.aux_doubles;
  double* Mt = reinterpret_cast<double*>(a.ws + a.L.off_m) + (size_t)s * a.L.m_doubles;
  double* Ut = aux + aux_ut(a.L);
  const int dk = a.L.dk;

  const double* h = a.hyp + (size_t)s * a.ndim;
  const double ss = h[0];
  const double noise = (a.ndim == d + 2) ? h[d + 1] * h[d + 1] : a.nugget_var;
  const double jit = a.jitter_s ? a.jitter_s[s] : a.jitter;
  const double SQRT2 = 1.4142135623730951, SQRTPI = 1.7724538509055159;
  if (tid < kCkPad) {
    const double c = (tid < d) ? 1.0 / (SQRT2 * h[1 + tid]) : 0.0;
    ck[tid] = c;
    cku[tid] = kKappa * c;
  }
  if (tid >= 32 && tid < 32 + kTab64) tab[tid - 32] = ss * exp2((double)(tid - 32) / 64.0);
  for (int I = tid; I < T; I += kPtThreads)
    for (int J = 0; J <= I; J++) tij[I * (I + 1) / 2 + J] = (unsigned short)((I << 8) | J);
  if (tid == 0) {
    // hyper-parameter sanity (same rules as k_prepare)
    int bad = !(ss > 9.5e-7 && ss < 1.0e6) || !(noise >= 0.0 && noise < 1.0e6);
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
  for (int idx = tid; idx < kTab64 * kTabRep; idx += kPtThreads) aux[aux_tabr() + idx] = tab[idx / kTabRep];

  // ---- prescaled, centred coordinates U (fragment-tiled) and -|U_j|^2: for the scoring kernel and for the Gram tiles ----
  for (int idx = tid; idx < npad * 4 * dk; idx += kPtThreads) {
    const int blk = idx >> 5, ln = idx & 31;
    const int R = blk / dk, i = blk - R * dk;
    const int j = 8 * R + (ln >> 2), k = 4 * i + (ln & 3);
    const double u = (j < n && k < d) ? (a.X[(size_t)j * d + k] - 0.5) * cku[k] : 0.0;
    Ut[idx] = u;
    if (u_smem) Us[idx] = u;
  }
  for (int j = tid; j < npad; j += kPtThreads) {  // fixed order over k
    double nun = 0.0;
    if (j < n)
      for (int k = 0; k < d; k++) {
        const double u = (a.X[(size_t)j * d + k] - 0.5) * cku[k];
        nun = fma(-u, u, nun);
      }
    aux[aux_nw() + 2 * j] = nun;
    nunv[j] = nun;
  }
  // ---- bordered row tile T: row 0 = e = xik(X) (or the quadrature override), row 1 = Y, rows 2..7 = 0 ----
  double* border = tiles + pt_slot(T, 0);
  for (int idx = tid; idx < T * 64; idx += kPtThreads) border[idx] = 0.0;
  __syncthreads();
  for (int j = tid; j < n; j += kPtThreads) {
    // _core.py:86-92: ss * prod_k sqrt(pi)/sqrt(2) * ell_k * (erf((1-x)/(sqrt2 ell)) - erf((0-x)/(sqrt2 ell)))
    double p = 1.0;
    for (int k = 0; k < d; k++) {
      const double x = a.X[(size_t)j * d + k];
      p *= (SQRTPI / SQRT2) * h[1 + k] * (erf((1.0 - x) * ck[k]) - erf((0.0 - x) * ck[k]));
    }
    double* bt = border + (j >> 3) * 64;
    bt[pt_pos(0, j & 7)] = a.e_ovr ? a.e_ovr[(size_t)s * n + j] : ss * p;
    bt[pt_pos(1, j & 7)] = a.Y[j];
  }
  TPT(1);
  // ---- A = K + (noise + jitter) I as tiles, on the tensor pipe like the scoring kernel's cross-covariance:
  //      z[i][j] = -|U_i|^2 - |U_j|^2 + 2 U_i . U_j  (dk DMMAs per tile; the A fragment of row tile I and the B fragment of
  //      row tile J are the same fragment-order loads), K = variance * 2^(z/64) by the shared table exp2.  K(X, X) and
  //      k(X, x) are thereby computed by the same arithmetic.
  __syncthreads();
  {
    const double* Ub = u_smem ? Us : Ut;
    for (int t = warp; t < nlow; t += kPtWarps) {
      const int I = tij[t] >> 8, J = tij[t] & 255;
      const int i = 8 * I + r, j0 = 8 * J + 2 * q;
      const double ni = nunv[i];
      double z0 = ni + nunv[j0], z1 = ni + nunv[j0 + 1];
      const double* ua = Ub + (size_t)I * dk * 32 + lane;
      const double* ub = Ub + (size_t)J * dk * 32 + lane;
      for (int ii = 0; ii < dk; ii++) pt_dmma(z0, z1, ua[ii * 32], 2.0 * ub[ii * 32]);
      double v0 = rbf_exp64<1>(fmin(z0, 0.0), tab), v1 = rbf_exp64<1>(fmin(z1, 0.0), tab);
      // diagonal (GPy: diag.add(Ky, variance + 1e-8)); padded rows / columns: identity
      if (i >= n || j0 >= n) v0 = 0.0;
      if (i >= n || j0 + 1 >= n) v1 = 0.0;
      if (i == j0) v0 = (i < n) ? ss + (noise + jit) : 1.0;
      if (i == j0 + 1) v1 = (i < n) ? ss + (noise + jit) : 1.0;
      *reinterpret_cast<double2*>(tiles + t * 64 + pt_pos(r, 2 * q)) = make_double2(v0, v1);
    }
  }
  __syncthreads();

  TPT(2);
  // ---- left-looking blocked Cholesky of the bordered matrix ------------------------------------------------------
  int info = fail;  // -1: rejected hyper-parameters
#ifdef BODE_TIMELINE
  long long tacc_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  for (int J = 0; J < T && info == 0; J++) {
    // this warp's tiles of block column J: warp 0 the diagonal tile, warp w >= 1 rows I = J + w + 7m (I <= T)
    double c0[kPtHeld], c1[kPtHeld];
    int Im[kPtHeld];
    int cnt = 0;
    if (warp == 0) {
      Im[0] = J;
      cnt = 1;
    } else {
#pragma unroll
      for (int m = 0; m < kPtHeld; m++) {
        const int I = J + warp + 7 * m;
        if (I <= T) { Im[m] = I; cnt = m + 1; }
      }
    }
#pragma unroll
    for (int m = 0; m < kPtHeld; m++) { c0[m] = 0.0; c1[m] = 0.0; if (m >= cnt) Im[m] = J; }
    TPT_ACC_BEGIN();
    if (cnt > 0) {
      const double* rowJ = tiles + pt_slot(J, 0) + lane;
      const double* rowm[kPtHeld];
#pragma unroll
      for (int m = 0; m < kPtHeld; m++) rowm[m] = tiles + pt_slot(Im[m], 0) + lane;
      if (cnt == 1) pt_update<1>(c0, c1, rowJ, rowm, J);
      else if (cnt == 2) pt_update<2>(c0, c1, rowJ, rowm, J);
      else if (cnt == 3) pt_update<3>(c0, c1, rowJ, rowm, J);
      else pt_update<4>(c0, c1, rowJ, rowm, J);
#pragma unroll
      for (int m = 0; m < kPtHeld; m++)
        if (m < cnt) {
          const double2 o = *reinterpret_cast<const double2*>(tiles + pt_slot(Im[m], J) + pt_pos(r, 2 * q));
          c0[m] = o.x - c0[m];
          c1[m] = o.y - c1[m];
        }
    }
    TPT_ACC(8, 1);
    TPT_ACC(13, 0);
    if (warp == 0) {
      // 8 x 8 Cholesky: lane c (mod 8) owns row c
      const int c = lane & 7;
      double B[8];
#pragma unroll
      for (int cc = 0; cc < 8; cc++) B[cc] = __shfl_sync(0xffffffffu, (cc & 1) ? c1[0] : c0[0], 4 * c + (cc >> 1));
      double rd[8], dg = 1.0;  // dg: this lane's diagonal entry L[c][c] (no dynamic register indexing: B stays in registers)
      int bad = 0;
#pragma unroll
      for (int cc = 0; cc < 8; cc++) {
        const double piv = __shfl_sync(0xffffffffu, B[cc], cc);
        if (!(piv > 0.0) && bad == 0) bad = 8 * J + cc + 1;
        double rs = rsqrt(piv);
        rs = fma(rs * 0.5, fma(-piv * rs, rs, 1.0), rs);  // one Newton step (measured: free, the step is latency-bound elsewhere)
        rd[cc] = rs;                                        // 1 / L[cc][cc]
        if (c == cc) { B[cc] = piv * rs; dg = B[cc]; }
        else if (c > cc) B[cc] = B[cc] * rs;
#pragma unroll
        for (int c2 = cc + 1; c2 < 8; c2++) {
          const double lc2 = __shfl_sync(0xffffffffu, B[cc], c2);  // l_{c2,cc}
          if (c2 <= c) B[c2] = fma(-B[cc], lc2, B[c2]);
        }
      }
      // W = L_JJ^-1 by forward substitution: lane c computes column c (x[i] = W[i][c])
      double x[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        double t = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; k++) {
          const double lik = __shfl_sync(0xffffffffu, B[k], i);  // L[i][k]
          if (k >= c) t = fma(-lik, x[k], t);
        }
        x[i] = (i >= c) ? t * rd[i] : 0.0;
      }
      double* dt = tiles + pt_slot(J, J);
      if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) dt[pt_pos(i, c)] = x[i];
        ldiag[8 * J + c] = dg;
      }
      if (lane == 0 && bad != 0 && bad <= n) fail = bad;
    }
    TPT_ACC(9, 0);
    __syncthreads();
    TPT_ACC(14, 1);
    if (fail) { info = fail; break; }
    if (warp != 0 && cnt > 0) {
      // L[I][J] = C W_J^T: A fragments of C by register shuffles, B fragments = W_J in fragment order
      const double* wt = tiles + pt_slot(J, J) + lane;
      const double w0 = wt[0], w1 = wt[32];
      const int src0 = (lane & ~3) | (q >> 1), src1 = src0 | 2;
#pragma unroll
      for (int m = 0; m < kPtHeld; m++)
        if (m < cnt) {
          const double p00 = __shfl_sync(0xffffffffu, c0[m], src0), p01 = __shfl_sync(0xffffffffu, c1[m], src0);
          const double p10 = __shfl_sync(0xffffffffu, c0[m], src1), p11 = __shfl_sync(0xffffffffu, c1[m], src1);
          const double a0 = (lane & 1) ? p01 : p00, a1 = (lane & 1) ? p11 : p10;
          double r0 = 0.0, r1 = 0.0;
          pt_dmma(r0, r1, a0, w0);
          pt_dmma(r0, r1, a1, w1);
          *reinterpret_cast<double2*>(tiles + pt_slot(Im[m], J) + pt_pos(r, 2 * q)) = make_double2(r0, r1);
        }
    }
    __syncthreads();
  }

  TPT(3);
  TPT_ACC_FLUSH();
  double* cst = aux;  // consts
  if (info) {
    // failed factorisation: poison the sample (NaN propagates to every EKLD of this sample), zero the factors.
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (tid == 0) {
      cst[C_SS] = ss; cst[C_NOISE] = noise; cst[C_SIGMA0] = qnan; cst[C_SIGMA1] = qnan; cst[C_MU1] = qnan;
      cst[C_LOGDET] = qnan; cst[C_QUAD] = qnan; cst[C_LOGLIK] = -INFINITY;
      a.info[s] = info;
    }
    for (int j = tid; j < npad; j += kPtThreads) aux[aux_nw() + 2 * j + 1] = 0.0;
    if (!a.loglik_only)
      for (size_t idx = tid; idx < a.L.m_doubles; idx += kPtThreads) Mt[idx] = 0.0;
    return;
  }

  // ---- a, z and the scalars: sigma0 (bk), sigma1, mu1, logdet, quad ---------------------------------------------
  for (int j = tid; j < npad; j += kPtThreads) {
    const double* bt = border + (j >> 3) * 64;
    av[j] = (j < n) ? bt[pt_pos(0, j & 7)] : 0.0;
    zv[j] = (j < n) ? bt[pt_pos(1, j & 7)] : 0.0;
  }
  __syncthreads();
  double paa = 0.0, pza = 0.0, pzz = 0.0, pld = 0.0;
  for (int j = tid; j < n; j += kPtThreads) {
    const double avj = av[j], zvj = zv[j];
    paa = fma(avj, avj, paa);
    pza = fma(zvj, avj, pza);
    pzz = fma(zvj, zvj, pzz);
    pld += log(ldiag[j]);
  }
  const double saa = pt_block_sum(paa, red);
  const double sza = pt_block_sum(pza, red);
  const double szz = pt_block_sum(pzz, red);
  const double sld = 2.0 * pt_block_sum(pld, red);
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

  TPT(4);
  // ---- M = L^-1, row tile by row tile, in place (slot (I, I) already holds W_I = M[I][I]) ---------------------------
  for (int I = 1; I < T; I++) {
    const double* wt = tiles + pt_slot(I, I) + lane;
    const double w0 = wt[0], w1 = wt[32];
    // N[I][K] = W_I L[I][K], K < I: B operand = the tile itself (rows 4h..4h+3), read before anything is overwritten
    {
      double n0[kPtHeld], n1[kPtHeld];
#pragma unroll
      for (int m = 0; m < kPtHeld; m++) {
        const int K = warp + kPtWarps * m;
        n0[m] = 0.0;
        n1[m] = 0.0;
        if (K < I) {
          const double* t = tiles + pt_slot(I, K);
          const double b0 = t[(r >> 2) * 32 + q * 4 + (r & 3)];        // element [q][r]
          const double b1 = t[(r >> 2) * 32 + (q + 4) * 4 + (r & 3)];  // element [q + 4][r]
          pt_dmma(n0[m], n1[m], w0, b0);
          pt_dmma(n0[m], n1[m], w1, b1);
        }
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < kPtHeld; m++) {
        const int K = warp + kPtWarps * m;
        if (K < I) *reinterpret_cast<double2*>(tiles + pt_slot(I, K) + pt_pos(r, 2 * q)) = make_double2(n0[m], n1[m]);
      }
    }
    __syncthreads();
    // M[I][J] = - sum_{K = J}^{I-1} N[I][K] M[K][J], J < I: this warp's columns J = warp + 8m
    double m0[kPtHeld], m1[kPtHeld], e0[kPtHeld], e1[kPtHeld];
#pragma unroll
    for (int m = 0; m < kPtHeld; m++) { m0[m] = 0.0; m1[m] = 0.0; e0[m] = 0.0; e1[m] = 0.0; }
    {
      const double* rowI = tiles + pt_slot(I, 0) + lane;
      const int boff = (r >> 2) * 32 + q * 4 + (r & 3);  // element [q][r] of a tile; [q + 4][r] is 16 words further
      // column J_m = warp + 8m joins at K = J_m: segments of K with 1, 2, 3, 4 live columns
      pt_mpass<1>(m0, m1, e0, e1, rowI, tiles, boff, warp, warp, min(I, warp + kPtWarps));
      pt_mpass<2>(m0, m1, e0, e1, rowI, tiles, boff, warp, warp + kPtWarps, min(I, warp + 2 * kPtWarps));
      pt_mpass<3>(m0, m1, e0, e1, rowI, tiles, boff, warp, warp + 2 * kPtWarps, min(I, warp + 3 * kPtWarps));
      pt_mpass<4>(m0, m1, e0, e1, rowI, tiles, boff, warp, warp + 3 * kPtWarps, I);
#pragma unroll
      for (int m = 0; m < kPtHeld; m++) { m0[m] += e0[m]; m1[m] += e1[m]; }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < kPtHeld; m++) {
      const int Jm = warp + kPtWarps * m;
      if (Jm < I) *reinterpret_cast<double2*>(tiles + pt_slot(I, Jm) + pt_pos(r, 2 * q)) = make_double2(-m0[m], -m1[m]);
    }
  }
  __syncthreads();

  TPT(5);
  // ---- w = L^-T a = M^T a: thread per column, four interleaved partial sums in a fixed order ----------------------------
  for (int j = tid; j < npad; j += kPtThreads) {
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
    if (j < n) {
      const int Jt = j >> 3, jc = j & 7;
      auto m_at = [&](int i) { return tiles[pt_slot(i >> 3, Jt) + pt_pos(i & 7, jc)]; };
      int i = j;
      for (; i + 3 < n; i += 4) {
        p0 = fma(m_at(i), av[i], p0);
        p1 = fma(m_at(i + 1), av[i + 1], p1);
        p2 = fma(m_at(i + 2), av[i + 2], p2);
        p3 = fma(m_at(i + 3), av[i + 3], p3);
      }
      if (i < n) p0 = fma(m_at(i), av[i], p0);
      if (i + 1 < n) p1 = fma(m_at(i + 1), av[i + 1], p1);
      if (i + 2 < n) p2 = fma(m_at(i + 2), av[i + 2], p2);
    }
    wv[j] = (p0 + p1) + (p2 + p3);
  }
  __syncthreads();
  for (int j = tid; j < npad; j += kPtThreads) aux[aux_nw() + 2 * j + 1] = (j < n) ? wv[j] : 0.0;

  TPT(6);
  // ---- copy out: half h of tile (I, J) is block (R = I, g = 2J + h) of the tiled L^-1 (layout.h) ----------------------
  {
    for (int t = warp; t < nlow; t += kPtWarps) {
      const int I = tij[t] >> 8, J = tij[t] & 255;
      const double* src = tiles + t * 64;
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        const int g = 2 * J + hh;
        const int i = 8 * I + r, j = 4 * g + q;
        double v = src[hh * 32 + lane];
        if (i >= n || j >= n || j > i) v = 0.0;
        Mt[m_goff(T, g) + (size_t)(I - J) * 32 + lane] = v;
      }
    }
  }
  __syncthreads();
  TPT(7);
}

size_t prepare_tiles_smem_bytes(const Layout& L) {
  const size_t ntiles = (size_t)L.nrt * (L.nrt + 1) / 2 + L.nrt;
  const size_t fact = ntiles * 64 + 5 * (size_t)L.npad + 2 * kCkPad + kTab64 + 8 + (((size_t)L.nrt * (L.nrt + 1) / 2 + 3) >> 2);
  const size_t coef = 8 * 64;
  return sizeof(double) * (fact > coef ? fact : coef);
}
// opt-in shared memory per CTA of the current device (232 448 B on B200), queried once per thread and device
size_t smem_optin_bytes() {
  static thread_local int dev_cached = -1;
  static thread_local size_t bytes = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != dev_cached) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    bytes = (size_t)v;
    dev_cached = dev;
  }
  return bytes;
}

static size_t prepare_tiles_xs_bytes(const Layout& L) { return sizeof(double) * (size_t)L.npad * 4 * L.dk; }

cudaError_t launch_prepare_tiles(const PrepArgs& a_in, cudaStream_t st) {
  PrepArgs a = a_in;
  size_t smem = prepare_tiles_smem_bytes(a.L);
  a.x_smem = smem + prepare_tiles_xs_bytes(a.L) + 64 <= smem_optin_bytes() ? 1 : 0;  // stage the scaled tiles when there is room
  if (a.x_smem) smem += prepare_tiles_xs_bytes(a.L);
  static thread_local int attr_dev = -1;
  static thread_local size_t attr_smem = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_dev || smem > attr_smem) {
    e = cudaFuncSetAttribute(k_prepare_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
    attr_smem = smem;
  }
  const int extra = a.loglik_only ? 0 : a.S * ((a.d + 3) / 4);
  k_prepare_tiles<<<a.S + extra, kPtThreads, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace bode

// K2+K3+K4 -- dense EKLD scoring of a candidate block under every hyper-sample (sm_100a).
//
// Replaces N_d * S evaluations of avg_kld_mean (reference bode/_sampler.py:464-481) and of the helpers it calls
// (get_params_2 :391-402, predict / kern.K inside GPy, xik _core.py:86-92), and the log / mean / var / argmax of
// mcmc_kld + optimize (:492, :659).  Per (candidate x, hyper-sample s):
//     k   = K_ARD-RBF(X, x; theta_s)            (K2, phase A: GEMM-form distance on the tensor pipe + exp2, never leaves smem)
//     xi  = xik(x)                               (k_ekld_reduce: even Chebyshev series of the d factors, erf fallback)
//     t   = L_s^-1 k ; sigma_x = ss - t.t        (K3, phase B: FP64 tensor-core DMMA.8x8x4 against the tiled L^-1)
//     v   = xi - w_s.k                           (K3, folded into phase A)
//     EKLD = log(sqrt(s1)/sqrt(s2)) + s2/(2 s1) + (v^2/s1/(sigma_x+noise))*0.5 - 0.5 ; logk = log(EKLD)   (K4)
//
// Work decomposition: one CTA = one work item = (tile of C candidates) x (chunk of SC hyper-samples).  Per sample the
// CTA runs phase A (all warps build Kc[npad][C] in shared memory: the distance part on the FP64 tensor pipe, the
// exponential on the FP64 pipe), phase B (each warp owns up to RT row tiles x CT column tiles of T = L^-1 Kc as DMMA
// accumulators; L^-1 streams through a shared-memory ring filled with 1-D bulk copies (TMA, cp.async.bulk + mbarrier)
// by a rotating duty warp), and a short combine that writes sigma_x and v.  The column groups of a CTA run these
// phases as independent pipelines (their own named barrier).  k_ekld_reduce then evaluates xik(x), v and the closed-form
// EKLD (K4), sums log EKLD over s in hyper-sample order (deterministic, independent of how candidates were tiled or
// sharded across GPUs), does the block argmax, and its last block to finish picks the winner.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "layout.h"
#include "rbf.h"
#include "score_args.h"
#include "tuning.h"

namespace bode {

// ----------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, 1-D bulk copy (TMA), DMMA
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// non-blocking poll (try_wait may suspend the thread for a hardware time limit; test_wait never does)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// volatile variant: keeps the issue order written in the source (independent accumulators interleaved) -- the
// scheduler otherwise chains the dependent DMMAs of one tile back to back
__device__ __forceinline__ void dmma884_ordered(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// shared-memory counter increment with acquire-release semantics (the plain atomicAdd is relaxed)
__device__ __forceinline__ int atom_add_acq_rel_smem(int* p, int v) {
  int old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
// order this thread's (acquired) view of generic-proxy accesses to shared memory before a following async-proxy (TMA) access
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NT>
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
// one lane of a converged warp (no thread-id arithmetic, no S2R in hot loops)
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
  return p != 0;
}

// Developer timeline (tools/ts_run.py builds an out-of-tree copy with -DBODE_TIMELINE): clock64 stamps of 8 work items
// x 16 warps x 16 samples x 8 events, plus the cycles each warp spent waiting on the ring per sample.
#ifdef BODE_TIMELINE
#ifndef BODE_TS_BASE
#define BODE_TS_BASE 2000
#endif
__device__ long long g_ts[8 * 16 * 16 * 8];
__device__ int g_dbg_a;  // phase-A decomposition switches of the developer build: 8 no DMMA, 16 no exp, 32 no Kc store, 64 no loads
cudaError_t debug_set_phase_a(int v) { return cudaMemcpyToSymbol(g_dbg_a, &v, sizeof(int)); }
__device__ long long g_tw[8 * 16 * 16 * 2];
__device__ long long g_ta[8 * 16 * 16 * 4];
#define TS_STAMP(ev) do { if (ts_on && lane == 0 && ls < 16) g_ts[((ts_item * 16 + warp) * 16 + ls) * 8 + (ev)] = clock64(); } while (0)
#define TS_WAIT_BEGIN() long long tw0_ = clock64()
#define TS_WAIT_END(k) do { tw_acc[k] += clock64() - tw0_; } while (0)
#define DBG_PARAM , int dbg, long long* tsa
#define DBG_ARG , dbg, tsa
#define TSA(i) do { if (tsa && (threadIdx.x & 31) == 0) tsa[i] = clock64(); } while (0)
cudaError_t debug_copy_timeline(long long* host /* 8*16*16*10 */) {
  cudaError_t e = cudaMemcpyFromSymbol(host, g_ts, sizeof(g_ts));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyFromSymbol(host + 8 * 16 * 16 * 8, g_tw, sizeof(g_tw));
  if (e != cudaSuccess) return e;
  return cudaMemcpyFromSymbol(host + 8 * 16 * 16 * 10, g_ta, sizeof(g_ta));
}
#else
#define TS_STAMP(ev) do { } while (0)
#define TS_WAIT_BEGIN() do { } while (0)
#define TS_WAIT_END(k) do { } while (0)
#define DBG_PARAM
#define DBG_ARG
#define TSA(i) do { } while (0)
#endif

// ----------------------------------------------------------------------------------------------------------------
// Phase A: the cross-covariance tile Kc = k(X, candidates of this CTA) on the FP64 tensor pipe.
//   z[j][c] = -|U_j|^2 - |y_c|^2 + 2 U_j . y_c   (scaled coordinates of layout.h / rbf.h; DK DMMAs per 8 x 8 tile, the
//             A fragments come pre-tiled from the prepare kernel, the B fragments live in registers for the sample)
//   k[j][c] = variance * 2^(z/64)                 (9 FP64 operations, table lookups on bank-staggered copies)
// The warps of a column group (the RG warps that later contract these columns) build its CT column tiles: warp rg owns
// column tile rg % CT and the row tiles R == rg / CT (mod RG / CT); partial w.k per candidate; the xik factors
// (closed form, 2d erf per candidate) are spread over all threads as (candidate, dimension slice).
//
// Kc block (g, column tile) holds the 4 x 8 values of k4 group g in B-fragment order for phase B: position
// kk + 4 * n with  n = q + 4e  for tile column 2q + e -- a fixed permutation of the 8 candidates of a tile that makes
// these stores conflict-free (the accumulator column n of phase B belongs to tile column 2 (n & 3) + (n >> 2)).
// ----------------------------------------------------------------------------------------------------------------
// first half of a batch: fragment loads and the distance DMMAs of NB row tiles -> z (two columns per lane and tile), w
template <int DK, int NB>
__device__ __forceinline__ void pa_load_dmma(const double2* __restrict__ nw, const double* __restrict__ ut_lane, int R, int rstride,
                                             int r, const double (&yb)[DK], const double (&nyn)[2], double (&acc)[NB][2],
                                             double (&wv)[NB] DBG_PARAM) {
  double af[NB][DK];
#pragma unroll
  for (int t = 0; t < NB; t++) {
    const int Rt = R + t * rstride;
#ifdef BODE_TIMELINE
    if (dbg & 64) { wv[t] = 1.0; acc[t][0] = nyn[0]; acc[t][1] = nyn[1]; for (int i = 0; i < DK; i++) af[t][i] = nyn[0]; continue; }
#endif
    const double2 v = nw[8 * Rt + r];
    wv[t] = v.y;
    acc[t][0] = v.x + nyn[0];
    acc[t][1] = v.x + nyn[1];
#pragma unroll
    for (int i = 0; i < DK; i++) af[t][i] = ut_lane[(size_t)(Rt * DK + i) * 32];
  }
#pragma unroll
  for (int i = 0; i < DK; i++)
#pragma unroll
    for (int t = 0; t < NB; t++) {
#ifdef BODE_TIMELINE
      if (dbg & 8) { acc[t][0] += af[t][i]; continue; }
#endif
      dmma884(acc[t][0], acc[t][1], af[t][i], yb[i]);
    }
}

// second half: k = variance 2^(z/64), the partial w.k and the stores into the Kc tile
template <int NB>
__device__ __forceinline__ void pa_exp_store(const double (&acc)[NB][2], const double (&wv)[NB], int R, int rstride,
                                             unsigned tab_lane, double (&wk)[2], double* __restrict__ kst, int kstride DBG_PARAM) {
  double kv[2 * NB];
#ifdef BODE_TIMELINE
  if (dbg & 16) { for (int i = 0; i < 2 * NB; i++) kv[i] = (&acc[0][0])[i]; } else
#endif
  rbf_exp64_n<2 * NB, kTabRep>(&acc[0][0], tab_lane, kv);
#pragma unroll
  for (int t = 0; t < NB; t++) {
    wk[0] = fma(wv[t], kv[2 * t], wk[0]);
    wk[1] = fma(wv[t], kv[2 * t + 1], wk[1]);
    double* dst = kst + (size_t)(R + t * rstride) * kstride;
#ifdef BODE_TIMELINE
    if (dbg & 32) continue;
#endif
    dst[0] = kv[2 * t];
    dst[16] = kv[2 * t + 1];
  }
}

template <int DK, int NB>
__device__ __forceinline__ void phase_a_batch(const double2* __restrict__ nw, const double* __restrict__ ut_lane, int R, int rstride,
                                              int r, const double (&yb)[DK], const double (&nyn)[2],
                                              unsigned tab_lane, double (&wk)[2], double* __restrict__ kst,
                                              int kstride DBG_PARAM) {
  double acc[NB][2], wv[NB];
  pa_load_dmma<DK, NB>(nw, ut_lane, R, rstride, r, yb, nyn, acc, wv DBG_ARG);
  pa_exp_store<NB>(acc, wv, R, rstride, tab_lane, wk, kst, kstride DBG_ARG);
}

template <int DK, int C, int RG, int CT>
__device__ __forceinline__ void phase_a(const double* __restrict__ aux, const double* __restrict__ nwp, const double* __restrict__ Ut,
                                        int nrt, const double* __restrict__ xt, double* __restrict__ Kc,
                                        double* __restrict__ pv, int rg, int cg, int lane DBG_PARAM) {
  constexpr int NCT = C / 8, NRS = RG / CT;
  static_assert(RG % CT == 0, "the warps of a column group must split evenly over its column tiles");
  const int ctile = cg * CT + rg % CT, rs = rg / CT;
  const int r = lane >> 2, q = lane & 3;
  const double* cks = aux + aux_cks();
  const unsigned tab_lane = smem_u32(aux + aux_tabr() + (lane & (kTabRep - 1)));  // the aux block is in shared memory
  // B fragments (dimension 4i + q of candidate 8 ctile + r), doubled scaled coordinates
  double yb[DK];
#pragma unroll
  for (int i = 0; i < DK; i++) yb[i] = (xt[(4 * i + q) * C + 8 * ctile + r] - 0.5) * cks[4 * i + q];
  // -|y|^2 of the two candidates this thread's accumulators belong to (tile columns 2q, 2q + 1), fixed order over k
  // (the padded dimensions k >= d have cks = 0 and add exact zeros)
  // (even and odd k in separate chains, added at the end: half the dependent latency of this once-per-sample setup)
  double nyn[2], nye[2] = {0.0, 0.0}, nyo[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 4 * DK; k += 2) {
    const double c2 = cks[k], c3 = cks[k + 1];
    const double t0 = (xt[k * C + 8 * ctile + 2 * q] - 0.5) * c2, t1 = (xt[k * C + 8 * ctile + 2 * q + 1] - 0.5) * c2;
    const double t2 = (xt[(k + 1) * C + 8 * ctile + 2 * q] - 0.5) * c3, t3 = (xt[(k + 1) * C + 8 * ctile + 2 * q + 1] - 0.5) * c3;
    nye[0] = fma(-t0, t0, nye[0]);
    nye[1] = fma(-t1, t1, nye[1]);
    nyo[0] = fma(-t2, t2, nyo[0]);
    nyo[1] = fma(-t3, t3, nyo[1]);
  }
  nyn[0] = (nye[0] + nyo[0]) * 0.25;
  nyn[1] = (nye[1] + nyo[1]) * 0.25;
  TSA(0);
  double wk[2] = {0.0, 0.0};
  const double2* nw = reinterpret_cast<const double2*>(nwp);
  const double* ut_lane = Ut + lane;
  // element (row r of tile R, tile column 2q + e) -> block (2R + (r >> 2), ctile), position (r & 3) + 4q + 16e
  double* kst = Kc + (size_t)((r >> 2) * NCT + ctile) * 32 + (r & 3) + 4 * q;
  constexpr int kstride = 2 * NCT * 32;
  // four row tiles (eight independent exponentials per thread) per step; measured on B200: 2, 3, 4 or 6 tiles per
  // step are within 0.6 % of each other, the loop is throughput-bound
#ifndef BODE_PA_NB
#define BODE_PA_NB 4
#endif
  constexpr int NB = BODE_PA_NB;
  int R = rs;
  // software pipeline over the batches: the fragment loads and distance DMMAs of batch i+1 are issued before the
  // exponentials of batch i, so that the 52-cycle DMMA latency and the shared-memory round trips of one batch hide behind
  // the FP64 work of the other (phase A was latency-bound: 57 % pipe-busy with both warps of a scheduler in it)
  if (R + (NB - 1) * NRS < nrt) {
    double za[NB][2], wa[NB];
    pa_load_dmma<DK, NB>(nw, ut_lane, R, NRS, r, yb, nyn, za, wa DBG_ARG);
    int Rn = R + NB * NRS;
    for (; Rn + (NB - 1) * NRS < nrt; Rn += NB * NRS) {
      double zb[NB][2], wb[NB];
      pa_load_dmma<DK, NB>(nw, ut_lane, Rn, NRS, r, yb, nyn, zb, wb DBG_ARG);
      pa_exp_store<NB>(za, wa, R, NRS, tab_lane, wk, kst, kstride DBG_ARG);
#pragma unroll
      for (int t = 0; t < NB; t++) { za[t][0] = zb[t][0]; za[t][1] = zb[t][1]; wa[t] = wb[t]; }
      R = Rn;
    }
    pa_exp_store<NB>(za, wa, R, NRS, tab_lane, wk, kst, kstride DBG_ARG);
    R = Rn;
  }
  int left = (R < nrt) ? (nrt - R + NRS - 1) / NRS : 0;
  if (NB > 4) {
    while (left >= 4) { phase_a_batch<DK, 4>(nw, ut_lane, R, NRS, r, yb, nyn, tab_lane, wk, kst, kstride DBG_ARG); R += 4 * NRS; left -= 4; }
  }
  if (left == 3) phase_a_batch<DK, 3>(nw, ut_lane, R, NRS, r, yb, nyn, tab_lane, wk, kst, kstride DBG_ARG);
  else if (left == 2) phase_a_batch<DK, 2>(nw, ut_lane, R, NRS, r, yb, nyn, tab_lane, wk, kst, kstride DBG_ARG);
  else if (left == 1) phase_a_batch<DK, 1>(nw, ut_lane, R, NRS, r, yb, nyn, tab_lane, wk, kst, kstride DBG_ARG);
  TSA(1);
  // partial w.k of this warp's rows: sum over the 8 row lanes (fixed tree), one slot per (row slice, candidate)
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    wk[0] += __shfl_xor_sync(0xffffffffu, wk[0], o);
    wk[1] += __shfl_xor_sync(0xffffffffu, wk[1], o);
  }
  if (lane < 4) {
    pv[rs * C + 8 * ctile + 2 * q] = wk[0];
    pv[rs * C + 8 * ctile + 2 * q + 1] = wk[1];
  }
  TSA(2);
}

template <int C, int RG, int CT>
__device__ __forceinline__ void phase_a_dispatch(const double* aux, const double* nwp, const double* Ut, int nrt, int dk,
                                                 const double* xt, double* Kc, double* pv, int rg, int cg, int lane DBG_PARAM) {
  switch (dk) {
    case 1: phase_a<1, C, RG, CT>(aux, nwp, Ut, nrt, xt, Kc, pv, rg, cg, lane DBG_ARG); break;
    case 2: phase_a<2, C, RG, CT>(aux, nwp, Ut, nrt, xt, Kc, pv, rg, cg, lane DBG_ARG); break;
    case 3: phase_a<3, C, RG, CT>(aux, nwp, Ut, nrt, xt, Kc, pv, rg, cg, lane DBG_ARG); break;
    default: phase_a<4, C, RG, CT>(aux, nwp, Ut, nrt, xt, Kc, pv, rg, cg, lane DBG_ARG); break;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Phase B primitive: one PAIR of k4 groups (8 columns of L^-1) against the accumulator slots T0..RT-1.
// mb points at block (row tile p, group 2p) of the staged chunk (+lane); toff[t] = 32 * (tile index of slot t);
// gstride = doubles between group 2p and group 2p+1.  All fragment loads are issued before the DMMAs.
// ----------------------------------------------------------------------------------------------------------------
template <int RT, int CT, int T0>
__device__ __forceinline__ void mma_pair(double (&acc)[RT][CT][2], const int (&toff)[RT], const double* __restrict__ kb,
                                         const double* __restrict__ m0, int gstride, int C) {
  double b0[CT], b1[CT], a0[RT - T0], a1[RT - T0];
#pragma unroll
  for (int ct = 0; ct < CT; ct++) { b0[ct] = kb[ct * 32]; b1[ct] = kb[C * 4 + ct * 32]; }
#pragma unroll
  for (int t = T0; t < RT; t++) { a0[t - T0] = m0[toff[t]]; a1[t - T0] = m0[gstride + toff[t]]; }
#pragma unroll
  for (int t = T0; t < RT; t++)
#pragma unroll
    for (int ct = 0; ct < CT; ct++) dmma884(acc[t][ct][0], acc[t][ct][1], a0[t - T0], b0[ct]);
#pragma unroll
  for (int t = T0; t < RT; t++)
#pragma unroll
    for (int ct = 0; ct < CT; ct++) dmma884(acc[t][ct][0], acc[t][ct][1], a1[t - T0], b1[ct]);
}

// ----------------------------------------------------------------------------------------------------------------
// The scoring kernel.  RG x CG warps = RG row groups x CG column groups; each warp owns up to RT row tiles
// (8 rows of T each) x CT column tiles (8 candidates each) as DMMA accumulators.  C = CG*CT*8 candidates per CTA.
// ----------------------------------------------------------------------------------------------------------------
// Static recursion over the number of dead accumulator slots: slot t dies at pair pend[t]; while slots T0.. are
// live the warp runs the T0-specialised pair primitive.  The chunk transition (release the finished ring stage,
// wait for the next) is an inline, rarely taken branch of the pair loop.
template <int RT, int CT, int T0>
struct PhaseB {
  template <typename Enter, typename Leave>
  static __device__ __forceinline__ void run(double (&acc)[RT][CT][2], const int (&toff)[RT], const int (&pend)[RT],
                                             int& p, int& next_cp, const double*& mb, const double*& kb, int nrt, int C,
                                             Enter&& enter, Leave&& leave) {
    const int pe = pend[T0];
    while (p < pe) {
      if (p == next_cp) { leave(); enter(); }
      const int gstride = (nrt - p) * 32;
      mma_pair<RT, CT, T0>(acc, toff, kb, mb - p * 32, gstride, C);
      mb += 2 * gstride;
      kb += 2 * C * 4;
      ++p;
    }
    PhaseB<RT, CT, T0 + 1>::run(acc, toff, pend, p, next_cp, mb, kb, nrt, C, enter, leave);
  }
};
template <int RT, int CT>
struct PhaseB<RT, CT, RT> {
  template <typename Enter, typename Leave>
  static __device__ __forceinline__ void run(double (&)[RT][CT][2], const int (&)[RT], const int (&)[RT], int&, int&,
                                             const double*&, const double*&, int, int, Enter&&, Leave&&) {}
};

template <int RG, int CG, int RT, int CT>
__global__ void __launch_bounds__(32 * RG * CG, (RG * CG <= 8 && RT * CT <= 16) ? 2 : 1) k_score(const __grid_constant__ ScoreArgs a) {
  constexpr int kWarps = RG * CG;        // 8 (255 registers/thread) or 16 (128), one CTA per SM
  constexpr int kThreads = 32 * kWarps;
  constexpr int C = CG * CT * 8;         // candidates per CTA
  constexpr int GC = CT * 8;             // candidates per column group
  constexpr int GT = RG * 32;            // threads per column group
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Layout& L = a.L;
  const int nrt = L.nrt, npad = L.npad;
  const int stages = a.stages, nchunks = a.nchunks;

  // ---- shared memory carve-up (offsets in doubles; everything 16-byte aligned) ----
  double* Kc = reinterpret_cast<double*>(smem_raw);             // [ng][C][4]
  double* auxs = Kc + (size_t)npad * C;                          // aux block of the current sample
  const int aux_smem_doubles = a.u_global ? aux_nw() : (int)L.aux_doubles;  // nw and Ut may stay in global memory
  double* ring = auxs + aux_smem_doubles;                        // [stages][chunk_max]
  double* red = ring + (size_t)stages * a.chunk_max;             // [RG][C]
  double* pv = red + RG * C;                                     // [2][NRS][C]   (NRS*C = 8 * kWarps)
  double* xt = pv + 2 * 8 * kWarps;                              // [4 dk][C]  candidate tile, dimension-major
  const int xt_rows = 4 * L.dk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xt + C * xt_rows);
  const uint32_t full0 = smem_u32(bars);                          // full[stages]: TMA completion per ring stage
  const uint32_t empty0 = full0 + 8 * kMaxStages;                 // empty[stages]: all warps are done with the stage
  const uint32_t aux_full = empty0 + 8 * kMaxStages;
  const uint32_t stagger = aux_full + 8;                          // first phase A of the even column groups is done
  int* aux_cnt = reinterpret_cast<int*>(bars + 2 * kMaxStages + 2);    // warps that finished phase A so far (all samples)
  int* chunk_off = aux_cnt + 2;                                         // [nchunks+1] offset (doubles) of chunk q inside M_s
  int* chunk_p0 = chunk_off + nchunks + 1;                              // [nchunks+2] first group PAIR of chunk q
  int* stage_cnt = chunk_p0 + nchunks + 2;                              // [stages] warps that left the stage so far (all uses)

  // ---- work item ----
  const int item = blockIdx.x;
  const int sch = item / a.ntiles, tile = item - sch * a.ntiles;
  // hyper-sample chunk of this item: the schedule is n_main equal chunks followed by a tapered tail (largest first, so
  // that the last wave of CTAs is made of short items)
  int s_begin, s_end;
  if (sch < a.n_main) {
    s_begin = sch * a.sc;
    s_end = min(a.S, s_begin + a.sc);
  } else {
    s_begin = a.tail_s0[sch - a.n_main];
    s_end = a.tail_s0[sch - a.n_main + 1];
  }
  const int nsamp = s_end - s_begin;
  const int64_t c0 = (int64_t)tile * C;

  if (tid == 0) {
    for (int i = 0; i < stages; i++) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, kWarps);
    }
    mbar_init(aux_full, 1);
    mbar_init(stagger, RG * ((CG + 1) / 2));
    aux_cnt[0] = 0;
    for (int i = 0; i < stages; i++) stage_cnt[i] = 0;
    if (blockIdx.x == 0) *a.done_counter = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int q = tid; q <= nchunks; q += kThreads) {
    chunk_off[q] = (int)m_goff(nrt, a.chunk_g0[q]);
    chunk_p0[q] = (q < nchunks) ? (a.chunk_g0[q] >> 1) : 0x7fffffff;
  }
  if (tid == 0) chunk_p0[nchunks + 1] = 0x7fffffff;  // read-ahead slot
  // candidate tile -> smem, dimension-major (rows beyond Nd are filled with a harmless interior point)
  for (int idx = tid; idx < C * xt_rows; idx += kThreads) {
    const int k = idx / C, c = idx - k * C;
    double v = 0.5;
    if (k < L.d && c0 + c < a.Nd) v = a.Xd[(c0 + c) * L.d + k];
    xt[idx] = v;
  }
  __syncthreads();

  const char* wsb = static_cast<const char*>(a.ws);
  const double* aux_g = reinterpret_cast<const double*>(wsb + L.off_aux) + (size_t)s_begin * L.aux_doubles;
  const double* m_g = reinterpret_cast<const double*>(wsb + L.off_m) + (size_t)s_begin * L.m_doubles;
  const uint32_t aux_bytes = (uint32_t)(aux_smem_doubles * sizeof(double));
  const uint32_t ring_u32 = smem_u32(ring), chunk_bytes_max = (uint32_t)a.chunk_max * 8u;

  // There is no producer warp (every warp computes).  Ring protocol per chunk `it`: every warp waits full[st],
  // computes, and arrives on empty[st] (fire and forget); the refill of the stage used by chunk it-lag is issued when
  // chunk it is entered, by the warp on refill duty (see `enter` below).
  auto issue_chunk = [&](int ls_, int q_, int st_) {
    const uint32_t bytes = (uint32_t)(chunk_off[q_ + 1] - chunk_off[q_]) * 8u;
    mbar_arrive_expect_tx(full0 + 8 * st_, bytes);
    bulk_copy_g2s(ring_u32 + (uint32_t)st_ * chunk_bytes_max, m_g + (size_t)ls_ * L.m_doubles + chunk_off[q_], bytes, full0 + 8 * st_);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(aux_full, aux_bytes);
    bulk_copy_g2s(smem_u32(auxs), aux_g, aux_bytes, aux_full);
    int ls_ = 0, q_ = 0;
    for (int i = 0; i < stages && ls_ < nsamp; i++) {
      issue_chunk(ls_, q_, i);
      if (++q_ == nchunks) { q_ = 0; ls_++; }
    }
  }

  // Column groups are independent pipelines: the RG warps of group cg build their own columns of Kc (phase A),
  // contract them (phase B) and synchronise only among themselves (named barrier 1 + cg); the groups share the L^-1
  // ring, the aux block and the FP64 pipe.  Warp w runs on sub-partition w % 4, which hosts one row group of every
  // column group, so the barrier bubbles of one group fill with the work of another (measured: -2.8 % at config 3
  // against CTA-wide barriers).  Starting the odd groups one phase A late on purpose (BODE_STAGGER=1) does not pay:
  // the two-stage ring lets a group lead by one chunk at most, and deeper rings cost more in chunk transitions.
  const int rg = warp % RG, cg = warp / RG;
  const int tg = rg * 32 + lane;  // thread index inside the column group
  const int cbase = cg * GC;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + cg), "n"(GT) : "memory"); };
  int toff[RT], pend[RT];
#pragma unroll
  for (int t = 0; t < RT; t++) {
    const int R = (int)a.tile_id[rg * kMaxRT + t];  // -1: unused slot
    toff[t] = 32 * R;
    pend[t] = R + 1;                                // slot t is live for pairs p <= R
  }
  const double* kc_lane = Kc + (size_t)cbase * 4 + lane;
  const double* ring_lane = ring + lane;

  const int aux_ut_off = aux_ut(L);
  int st = 0, it = 0, round_ = 1;  // round_: how many times the stages have been used, this use included
  uint32_t ph = 0;
  // refills trail the consumer by `lag` chunks so that the duty warp rarely has to wait for slower warps
  const int lag = (CG == 1 && stages >= 4) ? 2 : 1;  // (the lagging groups hold the duty when CG > 1: no extra lag needed)
  // refill cursor: (sample, chunk) that goes into the stage released by chunk it-lag, i.e. chunk it-lag+stages
  int rf_ls = 0, rf_q = 0;
  for (int i = 0; i < stages; i++) if (++rf_q == nchunks) { rf_q = 0; rf_ls++; }

#ifdef BODE_TIMELINE
  const int ts_item = (int)blockIdx.x - BODE_TS_BASE;
  const bool ts_on = ts_item >= 0 && ts_item < 8 && kWarps <= 16;
  const int dbg = g_dbg_a;
#endif
  for (int ls = 0; ls < nsamp; ls++) {
#ifdef BODE_TIMELINE
    long long tw_acc[2] = {0, 0};
#endif
    TS_STAMP(0);
    const int par = ls & 1;
    double* pvp = pv + par * 8 * kWarps;
    // ---- phase A ----
    mbar_wait(aux_full, (uint32_t)par);
    TS_STAMP(1);
#ifdef BODE_TIMELINE
    long long* tsa = (ts_on && ls < 16) ? &g_ta[((ts_item * 16 + warp) * 16 + ls) * 4] : nullptr;
#endif
    if (CG > 1 && ls == 0 && (cg & 1) && a.stagger) mbar_wait(stagger, 0);  // opt-in: odd groups one phase A behind
    double ss = 0;
    if (tg < GC) ss = auxs[C_SS];
    if (!((a.debug_skip & 1) && ls > 0)) {
      // two call sites so that the fragment loads stay address-space specific (LDS vs LDG) instead of generic
      if (a.u_global) {
        const double* ag = aux_g + (size_t)ls * L.aux_doubles;
        phase_a_dispatch<C, RG, CT>(auxs, ag + aux_nw(), ag + aux_ut_off, nrt, L.dk, xt, Kc, pvp, rg, cg, lane DBG_ARG);
      } else {
        phase_a_dispatch<C, RG, CT>(auxs, auxs + aux_nw(), auxs + aux_ut_off, nrt, L.dk, xt, Kc, pvp, rg, cg, lane DBG_ARG);
      }
      TS_STAMP(5);  // (developer timeline: slot 5 = cross-covariance done)
    }
    TS_STAMP(2);
    // This warp is done with the aux block.  The last warp of the CTA to get here fetches the next sample's block
    // (the groups run out of step, so nobody waits for it: a counter in shared memory instead of a barrier).
    __syncwarp();
    if (elect_one()) {
      if (CG > 1 && ls == 0 && !(cg & 1)) mbar_arrive(stagger);
      // release this warp's reads of the aux block (ordered before the elected lane by the __syncwarp above), acquire
      // everybody else's; the last arriver then crosses to the async proxy before the bulk copy overwrites the block
      const int done = atom_add_acq_rel_smem(aux_cnt, 1) + 1;
      if (done == kWarps * (ls + 1) && ls + 1 < nsamp) {
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(aux_full, aux_bytes);
        bulk_copy_g2s(smem_u32(auxs), aux_g + (size_t)(ls + 1) * L.aux_doubles, aux_bytes, aux_full);
      }
    }
    group_sync();
    TS_STAMP(3);

    // ---- phase B: T = L^-1 Kc on the FP64 tensor pipe ----
    double acc[RT][CT][2];
#pragma unroll
    for (int t = 0; t < RT; t++)
#pragma unroll
      for (int ct = 0; ct < CT; ct++) { acc[t][ct][0] = 0.0; acc[t][ct][1] = 0.0; }

    int q = -1, p = 0, next_cp = 0, next_cp_pref = chunk_p0[1];
    const double* mb = ring_lane;
    const double* kb = kc_lane;
    auto enter = [&]() {  // start chunk q+1 (ring stage st)
      ++q;
      // refill duty for the stage released `lag` chunks ago (chunk it-lag -> chunk it-lag+stages) rotates over the
      // warps of the odd column groups: they run one phase A behind, leave every chunk last and so rarely wait for
      // the stage's empty barrier, and the leading groups never block on a straggler
      constexpr int kDutyWarps = (CG > 1) ? RG * (CG / 2) : kWarps;
      const int dj = (it - lag) & (kDutyWarps - 1);
      const int duty_warp = (CG > 1) ? ((2 * (dj / RG) + 1) * RG + dj % RG) : dj;
      if (!a.refill_last && it >= lag && duty_warp == warp) {
        if (rf_ls < nsamp && !(a.debug_skip & 4) && elect_one()) {
          const int pst = (st >= lag) ? st - lag : st - lag + stages;
          TS_WAIT_BEGIN();
          mbar_wait(empty0 + 8 * pst, (st >= lag) ? ph : (ph ^ 1));
          TS_WAIT_END(1);
          issue_chunk(rf_ls, rf_q, pst);
        }
        __syncwarp();
      }
      if (!a.refill_last && it >= lag) { if (++rf_q == nchunks) { rf_q = 0; rf_ls++; } }
      next_cp = next_cp_pref;               // prefetched one transition ago (its shared-memory latency is off the path)
      next_cp_pref = chunk_p0[q + 2];
      TS_WAIT_BEGIN();
      if (!(a.debug_skip & 4)) mbar_wait(full0 + 8 * st, ph);
      TS_WAIT_END(0);
      mb = ring_lane + (size_t)st * a.chunk_max;
    };
    auto leave = [&]() {  // done with chunk q
      __syncwarp();
      if (a.refill_last) {
        // The last warp to leave the stage refills it (chunk it + stages): a counter instead of the empty barrier, nobody
        // waits.  acq_rel orders every warp's reads of the stage before the last arriver, which then crosses to the async
        // proxy before the bulk copy overwrites it (same scheme as the aux block).
        if (lane == 0) {
          const int done = atom_add_acq_rel_smem(stage_cnt + st, 1) + 1;
          if (done == kWarps * round_ && rf_ls < nsamp && !(a.debug_skip & 4)) {
            fence_proxy_async_smem();
            issue_chunk(rf_ls, rf_q, st);
          }
        }
        if (++rf_q == nchunks) { rf_q = 0; rf_ls++; }
      } else if (lane == 0) {
        mbar_arrive(empty0 + 8 * st);
      }
      if (++st == stages) { st = 0; ph ^= 1; ++round_; }
      ++it;
    };
    enter();
    if (!(a.debug_skip & 2)) PhaseB<RT, CT, 0>::run(acc, toff, pend, p, next_cp, mb, kb, nrt, C, enter, leave);
    while (q + 1 < nchunks) { leave(); enter(); }  // chunks past this warp's last live row tile
    leave();
    TS_STAMP(4);

    // ---- column sums of squares over this warp's rows -> red[rg][c] ----
#pragma unroll
    for (int ct = 0; ct < CT; ct++) {
      double p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int t = 0; t < RT; t++) {
        p0 = fma(acc[t][ct][0], acc[t][ct][0], p0);
        p1 = fma(acc[t][ct][1], acc[t][ct][1], p1);
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, o);
        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
      }
      if (lane < 4) {
        // accumulator columns n = 2 lane, 2 lane + 1 hold tile columns 2 (n & 3) + (n >> 2) (phase A's store order)
        const int col0 = (lane < 2) ? 4 * lane : 4 * lane - 7;
        red[rg * C + cbase + ct * 8 + col0] = p0;
        red[rg * C + cbase + ct * 8 + col0 + 2] = p1;
      }
    }
    group_sync();
    TS_STAMP(6);

    // ---- per-candidate combine: sigma_x = ss - |L^-1 k|^2 and w.k.  The closed-form EKLD itself (two logs, two square
    // roots, four divisions: ~2.8k cycles of dependent FP64 latency per candidate) is finished by k_ekld_reduce: evaluated
    // here it sat on this column group's critical path once per sample and cost k_score 7 % (measured, round 2).
    if (tg < GC) {
      const int cc = cbase + tg;  // this group's candidates
      double tt = red[cc];
#pragma unroll
      for (int r = 1; r < RG; r++) tt += red[r * C + cc];
      constexpr int NRS = RG / CT;
      double wk = 0.0;
#pragma unroll
      for (int j = 0; j < NRS; j++) wk += pvp[j * C + cc];
      if (c0 + cc < a.Nd) {
        const size_t o = (size_t)(s_begin + ls) * a.Nd + c0 + cc;
        a.sx[o] = ss - tt;                          // latent predictive variance (GPy predict, full_cov 1x1; _sampler.py:475)
        if (a.mode != kModeVariance) a.wk[o] = wk;  // w.k = e_k^T A^-1 k(X, x), the data term of v (_sampler.py:399)
      }
    }
    TS_STAMP(7);
#ifdef BODE_TIMELINE
    if (ts_on && lane == 0 && ls < 16) { g_tw[((ts_item * 16 + warp) * 16 + ls) * 2] = tw_acc[0]; g_tw[((ts_item * 16 + warp) * 16 + ls) * 2 + 1] = tw_acc[1]; }
#endif
  }
}

// ----------------------------------------------------------------------------------------------------------------
// K4 + reduction, ONE kernel: xik(x), v, the closed-form Gaussian EKLD and its log for every (hyper-sample, candidate),
// mean / variance over the hyper-samples, block argmax and (last block to finish) the final argmax.
// (xik reference bode/_core.py:86-92; get_params_2 _sampler.py:399-401; avg_kld_mean :474-481; mcmc_kld :492; argmax :659)
//
// A block owns 64 candidates (lane = candidate, two per lane); its eight warps take eight consecutive hyper-samples at
// a time -- one sample per warp, so the series lengths and table reads are warp-uniform -- and hand their log EKLD
// values to warp 0 through shared memory, which adds them to the running sums in hyper-sample order: the sum over
// hyper-samples is exactly the sequential s = 0..S-1 sum the oracle defines, independent of tiling, sharding and
// launch geometry.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kErWarps = 8;               // hyper-samples in flight per block

__device__ __forceinline__ bool better(double av, int64_t ai, double bv, int64_t bi) {
  // np.argmax semantics: NaN counts as the maximum; ties -> lowest index
  const bool an = isnan(av), bn = isnan(bv);
  if (an || bn) {
    if (an && bn) return ai < bi;
    return an;
  }
  if (av > bv) return true;
  if (av < bv) return false;
  return ai < bi;
}

// block winner -> thread 0 (bv, bi); sv / si: shared scratch of 8 entries each
__device__ __forceinline__ void block_best(double& bv, int64_t& bi, double* sv, int64_t* si) {
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int w = 1; w < 8; w++)
      if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
}

// the last block to finish combines the per-block winners (the comparison is a total order, so the result does not
// depend on which block runs this or in which order the partials are visited)
__device__ __forceinline__ void final_argmax(double bv, int64_t bi, double* part_val, int64_t* part_idx,
                                             unsigned int* done_counter, double* best, int64_t* best_idx, double* sv,
                                             int64_t* si, bool* last) {
  if (threadIdx.x == 0) {
    part_val[blockIdx.x] = bv;
    part_idx[blockIdx.x] = bi;
    __threadfence();
    *last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!*last) return;
  __threadfence();
  bv = -INFINITY;
  bi = INT64_MAX;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    const double pv_ = __ldcg(part_val + i);
    const int64_t pi_ = __ldcg(part_idx + i);
    if (better(pv_, pi_, bv, bi)) { bv = pv_; bi = pi_; }
  }
  __syncthreads();
  block_best(bv, bi, sv, si);
  if (threadIdx.x == 0) {
    best[0] = bv;
    best_idx[0] = bi;
    *done_counter = 0u;
  }
}

// kErCpl: candidates per lane (independent dependency chains per thread; 1 for small designs so that the grid still covers the SMs)
template <int kErCpl>
__global__ void __launch_bounds__(32 * kErWarps) k_ekld_reduce(const char* __restrict__ ws, Layout L, const double* __restrict__ Xd,
                                                               const double* __restrict__ xi_ovr, double* __restrict__ sx,
                                                               const double* __restrict__ wkplane, int S, int64_t Nd,
                                                               int64_t idx_offset, int form, int use_series, int keep_logk,
                                                               double* __restrict__ score, double* __restrict__ var,
                                                               double* __restrict__ part_val, int64_t* __restrict__ part_idx,
                                                               unsigned int* __restrict__ done_counter,
                                                               double* __restrict__ best, int64_t* __restrict__ best_idx) {
  constexpr int kErCand = 32 * kErCpl;              // candidates per block
  extern __shared__ __align__(16) double xc_all[];  // [kErWarps][d][kXcStride]: series of the xik factors, one sample per warp
  __shared__ double lkbuf[2][kErWarps][kErCand];    // log EKLD (then squared deviations) on their way to warp 0
  __shared__ double meanbuf[kErCand];
  __shared__ double sv[8];
  __shared__ int64_t si[8];
  __shared__ bool last;
  const int d = L.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* xc = xc_all + (size_t)warp * d * kXcStride;
  int64_t c[kErCpl], cq[kErCpl];
  double acc[kErCpl];
#pragma unroll
  for (int q = 0; q < kErCpl; q++) {
    c[q] = (int64_t)blockIdx.x * kErCand + 32 * q + lane;
    cq[q] = c[q] < Nd ? c[q] : Nd - 1;  // candidates beyond Nd are evaluated on a clamped index and never stored
    acc[q] = 0.0;
  }
  const int niter = (S + kErWarps - 1) / kErWarps;
  for (int it = 0; it < niter; it++) {
    const int s = it * kErWarps + warp;
    double lk[kErCpl];
#pragma unroll
    for (int q = 0; q < kErCpl; q++) lk[q] = 0.0;
    if (s < S) {
      const double* aux = reinterpret_cast<const double*>(ws + L.off_aux) + (size_t)s * L.aux_doubles;
      const double ss = aux[C_SS], noise = aux[C_NOISE], sigma1 = aux[C_SIGMA1];
      size_t o[kErCpl];
      double xi[kErCpl];
#pragma unroll
      for (int q = 0; q < kErCpl; q++) o[q] = (size_t)s * Nd + cq[q];
      if (xi_ovr) {
#pragma unroll
        for (int q = 0; q < kErCpl; q++) xi[q] = xi_ovr[o[q]];  // quadrature estimate (k_rowmean)
      } else {
        // this warp's sample: the live part of each series, [n_terms, a_0/2, a_1 .. a_(n-1)]
        __syncwarp();
        const double* src0 = reinterpret_cast<const double*>(ws + L.off_xc) + (size_t)s * L.xc_doubles;
        for (int k = 0; k < d; k++) {
          const double* src = src0 + (size_t)k * kXcStride;
          const int n = (int)src[0];
          if (lane <= n) xc[k * kXcStride + lane] = src[lane];
          if (lane + 32 <= n) xc[k * kXcStride + lane + 32] = src[lane + 32];
        }
        __syncwarp();
        // _core.py:86-92: ss * prod_k sqrt(pi)/sqrt(2) * ell_k * (erf((1-x)/(sqrt2 ell)) - erf((0-x)/(sqrt2 ell))), k ascending;
        // each factor from its even Chebyshev series (Clenshaw), or from erf where the series did not converge; the
        // candidates of this thread advance together (independent dependency chains)
        double xf[kErCpl];
#pragma unroll
        for (int q = 0; q < kErCpl; q++) xf[q] = 1.0;
        for (int k = 0; k < d; k++) {
          const double* a = xc + k * kXcStride;
          const int n = use_series ? (int)a[0] : 0;
          double xv[kErCpl], fk[kErCpl];
#pragma unroll
          for (int q = 0; q < kErCpl; q++) xv[q] = Xd[cq[q] * d + k];
          if (n > 0) {
            double w[kErCpl], w2[kErCpl], b1[kErCpl], b2[kErCpl];
#pragma unroll
            for (int q = 0; q < kErCpl; q++) {
              const double u = fma(2.0, xv[q], -1.0);
              w[q] = fma(2.0 * u, u, -1.0);  // w = 2 u^2 - 1
              w2[q] = 2.0 * w[q];
              b1[q] = 0.0;
              b2[q] = 0.0;
            }
#pragma unroll 4
            for (int j = n - 1; j >= 1; j--) {
              const double aj = a[1 + j];
#pragma unroll
              for (int q = 0; q < kErCpl; q++) {
                const double t = fma(w2[q], b1[q], aj) - b2[q];
                b2[q] = b1[q];
                b1[q] = t;
              }
            }
#pragma unroll
            for (int q = 0; q < kErCpl; q++) fk[q] = fma(w[q], b1[q], a[1]) - b2[q];
          } else {
            const double ck = aux[aux_ck() + k], ell = aux[aux_ell() + k];
#pragma unroll
            for (int q = 0; q < kErCpl; q++) fk[q] = 1.2533141373155001 * ell * (erf((1.0 - xv[q]) * ck) - erf((0.0 - xv[q]) * ck));
          }
#pragma unroll
          for (int q = 0; q < kErCpl; q++) xf[q] *= fk[q];
        }
#pragma unroll
        for (int q = 0; q < kErCpl; q++) xi[q] = ss * xf[q];
      }
#pragma unroll
      for (int q = 0; q < kErCpl; q++) {
        const double sigma_x = sx[o[q]];
        const double v = xi[q] - wkplane[o[q]];  // _sampler.py:399
        const double v2 = v * v;                 // :401
        const double kxx = sigma_x + noise;      // :397 predict(include_likelihood=True)
        const double sigma2 = sigma1 - v2 / kxx; // :400
        double kld;
        if (form == kFormLog1p) {
          kld = -0.5 * log1p(-(v2 / (sigma1 * kxx)));
        } else {
          // :480, same operation order as the reference expression
          kld = ((log(sqrt(sigma1) / sqrt(sigma2)) + (sigma2 / (2.0 * sigma1))) + ((v2 / sigma1) / (sigma_x + noise)) * 0.5) - 0.5;
        }
        lk[q] = log(kld);  // :492 np.log(kld_j)
        if (keep_logk && c[q] < Nd) sx[o[q]] = lk[q];
      }
    }
#pragma unroll
    for (int q = 0; q < kErCpl; q++) lkbuf[it & 1][warp][32 * q + lane] = lk[q];
    __syncthreads();
    // warp 0 adds the eight samples' values in hyper-sample order (the other warps move on: the buffers alternate, and
    // a buffer is rewritten only after the next barrier, which warp 0 reaches after these reads)
    if (warp == 0) {
#pragma unroll
      for (int j = 0; j < kErWarps; j++)
        if (it * kErWarps + j < S) {
#pragma unroll
          for (int q = 0; q < kErCpl; q++) acc[q] += lkbuf[it & 1][j][32 * q + lane];
        }
    }
  }
  double mean[kErCpl], a2[kErCpl];
#pragma unroll
  for (int q = 0; q < kErCpl; q++) {
    mean[q] = acc[q] / (double)S;
    a2[q] = 0.0;
  }
  if (var != nullptr) {
    // population variance of the logs about the mean: every warp re-reads its own samples' stores, warp 0 adds the
    // squares in hyper-sample order (same scheme as the mean)
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int q = 0; q < kErCpl; q++) meanbuf[32 * q + lane] = mean[q];
    }
    __syncthreads();
    for (int it = 0; it < niter; it++) {
      const int s = it * kErWarps + warp;
#pragma unroll
      for (int q = 0; q < kErCpl; q++) {
        double dl2 = 0.0;
        if (s < S && c[q] < Nd) {
          const double dl = sx[(size_t)s * Nd + c[q]] - meanbuf[32 * q + lane];
          dl2 = dl * dl;
        }
        lkbuf[it & 1][warp][32 * q + lane] = dl2;
      }
      __syncthreads();
      if (warp == 0) {
#pragma unroll
        for (int j = 0; j < kErWarps; j++)
          if (it * kErWarps + j < S) {
#pragma unroll
            for (int q = 0; q < kErCpl; q++) a2[q] += lkbuf[it & 1][j][32 * q + lane];
          }
      }
    }
  }
  double bv = -INFINITY;
  int64_t bi = INT64_MAX;
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < kErCpl; q++) {
      if (c[q] < Nd) {
        if (var != nullptr) var[c[q]] = a2[q] / (double)S;
        score[c[q]] = mean[q];
        if (better(mean[q], idx_offset + c[q], bv, bi)) { bv = mean[q]; bi = idx_offset + c[q]; }
      }
    }
  }
  block_best(bv, bi, sv, si);
  final_argmax(bv, bi, part_val, part_idx, done_counter, best, best_idx, sv, si, &last);
}

// ----------------------------------------------------------------------------------------------------------------
// Plain reduction over hyper-samples (sequential in s) + argmax: the uncertainty-sampling comparator's mean variance
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ logk, int S, int64_t Nd, int64_t idx_offset,
                                                double* __restrict__ score, double* __restrict__ var,
                                                double* __restrict__ part_val, int64_t* __restrict__ part_idx,
                                                unsigned int* __restrict__ done_counter, double* __restrict__ best,
                                                int64_t* __restrict__ best_idx) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double bv = -INFINITY;
  int64_t bi = INT64_MAX;
  if (c < Nd) {
    double acc = 0.0;
    for (int s = 0; s < S; s++) acc += logk[(size_t)s * Nd + c];
    const double mean = acc / (double)S;
    score[c] = mean;
    if (var != nullptr) {
      double a2 = 0.0;
      for (int s = 0; s < S; s++) {
        const double dl = logk[(size_t)s * Nd + c] - mean;
        a2 += dl * dl;
      }
      var[c] = a2 / (double)S;
    }
    bv = mean;
    bi = idx_offset + c;
  }
  __shared__ double sv[8];
  __shared__ int64_t si[8];
  __shared__ bool last;
  block_best(bv, bi, sv, si);
  final_argmax(bv, bi, part_val, part_idx, done_counter, best, best_idx, sv, si, &last);
}

// logk -> EKLD values (for callers that want the (S, Nd) matrix the reference would have produced)
__global__ void k_exp_inplace(double* __restrict__ v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = exp(v[i]);
}

// ----------------------------------------------------------------------------------------------------------------
// Host-side planning and launch
// ----------------------------------------------------------------------------------------------------------------
static size_t score_smem_bytes(const Layout& L, int C, int RG, int nthreads, int stages, int chunk_max, int nchunks, int u_global) {
  size_t dbl = (size_t)L.npad * C + (u_global ? (size_t)aux_nw() : L.aux_doubles) + (size_t)stages * chunk_max + (size_t)RG * C + (size_t)nthreads / 2 +
               (size_t)C * 4 * L.dk;
  return dbl * sizeof(double) + sizeof(uint64_t) * (2 * kMaxStages + 3) +
         sizeof(int) * (2 * (nchunks + 2) + 8 + kMaxStages) + 128;
}

// Longest-processing-time assignment of row tiles (weight R+1 ~ number of live k4 groups) to the row groups.
// Inside a group the tiles are stored ascending and right-aligned: slots [RT-cnt, RT), unused slots = -1.
static void assign_tiles(int nrt, int nrg, int rt_cap, short* tile_id /*[kMaxRG][kMaxRT]*/) {
  int load[kMaxRG], cnt[kMaxRG];
  short tmp[kMaxRG][kMaxRT];
  for (int g = 0; g < nrg; g++) { load[g] = 0; cnt[g] = 0; }
  for (int R = nrt - 1; R >= 0; R--) {
    int best = -1;
    for (int g = 0; g < nrg; g++)
      if (cnt[g] < rt_cap && (best < 0 || load[g] < load[best])) best = g;
    tmp[best][cnt[best]++] = (short)R;  // descending R
    load[best] += R + 1;
  }
  for (int g = 0; g < kMaxRG; g++)
    for (int t = 0; t < kMaxRT; t++) tile_id[g * kMaxRT + t] = -1;
  for (int g = 0; g < nrg; g++)
    for (int t = 0; t < cnt[g]; t++) tile_id[g * kMaxRT + (rt_cap - 1 - t)] = tmp[g][t];
}

static int plan_chunks(const Layout& L, size_t smem_limit, size_t chunk_max, ScorePlan* P) {
  const int nrt = L.nrt;
  const int C = P->C;
  // chunk plan: consecutive PAIRS of k4 groups (both groups of a pair share their first stored row tile), greedily
  // packed up to the chunk size
  int nch = 0;
  int g = 0;
  P->chunk_g0[0] = 0;
  while (g < L.ng) {
    size_t acc = 0;
    int g1 = g;
    while (g1 < L.ng) {
      const size_t gs = (size_t)(nrt - (g1 >> 1)) * 64;
      if (acc + gs > chunk_max && g1 > g) break;
      acc += gs;
      g1 += 2;
    }
    g = g1;
    P->chunk_g0[++nch] = (unsigned short)g;
    if (nch >= kMaxChunks) { if (g < L.ng) return -1; break; }
  }
  P->nchunks = nch;
  P->chunk_max = (int)chunk_max;
  const int nthreads = 32 * P->rg * P->cg;
  int max_stages = 2, u_global = 0;  // two stages measured best (deeper rings mean smaller chunks, more transitions)
  if (tuning().stages > 0) max_stages = tuning().stages < 2 ? 2 : (tuning().stages > kMaxStages ? kMaxStages : tuning().stages);
  int stages = max_stages;
  while (stages >= 2 && score_smem_bytes(L, C, P->rg, nthreads, stages, (int)chunk_max, nch, 0) > smem_limit) stages--;
  if (stages < 2) {  // large n: leave the prescaled observations in global memory (read through L1/L2 in phase A)
    u_global = 1;
    stages = max_stages;
    while (stages >= 2 && score_smem_bytes(L, C, P->rg, nthreads, stages, (int)chunk_max, nch, 1) > smem_limit) stages--;
    if (stages < 2) return -1;
  }
  if (stages > nch + 1) stages = nch + 1;
  if (stages < 2) stages = 2;
  P->u_global = u_global;
  P->stages = stages;
  P->smem = score_smem_bytes(L, C, P->rg, nthreads, stages, (int)chunk_max, nch, u_global);
  return 0;
}

static int plan_variant(const Layout& L, size_t smem_limit, ScorePlan* P) {
  P->C = P->cg * P->ct * 8;
  const size_t pair0_doubles = (size_t)L.nrt * 64;
  if (tuning().chunk_kb > 0) {  // tuning knob
    size_t cm = (size_t)tuning().chunk_kb * 128;
    return plan_chunks(L, smem_limit, cm < pair0_doubles ? pair0_doubles : cm, P);
  }
  // Every chunk transition costs each warp ~450 cycles of barrier bookkeeping and two ring stages measured best, so:
  // the smallest number k of chunks whose balanced partition (minimal largest chunk, pairs of k4 groups as atoms)
  // fits twice in shared memory; the observations stay in shared memory unless no k allows it (large n).
  auto balanced_max = [&](int k) -> size_t {  // minimal largest chunk (doubles) of a split into <= k contiguous chunks
    size_t lo = pair0_doubles, hi = L.m_doubles;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      int cnt = 1;
      size_t acc = 0;
      for (int p = 0; p < L.nrt; p++) {
        const size_t ps = (size_t)(L.nrt - p) * 64;
        if (acc + ps > mid) { cnt++; acc = 0; }
        acc += ps;
      }
      if (cnt <= k) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  ScorePlan best = *P;
  bool have = false;
  const int kmax = L.nrt < kMaxChunks ? L.nrt : kMaxChunks;
  for (int k = 1; k <= kmax; k++) {
    ScorePlan T = *P;
    if (plan_chunks(L, smem_limit, balanced_max(k), &T) != 0) continue;
    if (!have || (best.u_global && !T.u_global)) { best = T; have = true; }
    if (!T.u_global) break;  // fewest chunks that keep the observations in shared memory
  }
  if (!have) return -1;
  *P = best;
  return 0;
}

int plan_score(const Layout& L, int64_t Nd, int sm_count, size_t smem_optin, ScorePlan* P) {
  const int nrt = L.nrt;
  // BODE_SCORE_VARIANT=<id> forces one configuration (tuning knob); the default order is the measured
  // preference: 8 warps with 32 accumulator tiles each (255 registers) beat 16 warps with 16 tiles (128 registers)
  const int forced = tuning().score_variant;
  struct Cand { int variant, rg, cg, rt, ct; size_t limit; bool only_forced; };
  const Cand cands[] = {
      {3, 4, 2, 8, 2, (228 * 1024) / 2 - 1024, true},  // two CTAs per SM, 32 candidates each (slower: L^-1 is streamed twice)
      {4, 4, 2, 8, 4, smem_optin, false},    // n <= 256 : 64 candidates per CTA,  8 warps x (8 x 4) tiles
      {0, 4, 4, 8, 2, smem_optin, false},    // n <= 256 : 64 candidates per CTA, 16 warps x (8 x 2) tiles
      {6, 4, 2, 16, 2, smem_optin, false},   // n <= 512 : 32 candidates per CTA,  8 warps x (16 x 2) tiles
      {1, 4, 4, 16, 1, smem_optin, false},   // n <= 512 : 32 candidates per CTA, 16 warps x (16 x 1) tiles
      {7, 8, 1, 16, 2, smem_optin, false},   // n <= 1024: 16 candidates per CTA,  8 warps x (16 x 2) tiles
      {2, 8, 2, 16, 1, smem_optin, false},   // n <= 1024: 16 candidates per CTA, 16 warps x (16 x 1) tiles
      {5, 16, 1, 8, 1, smem_optin, false},   // n <= 1024:  8 candidates per CTA (largest L^-1 chunks)
  };
  bool ok = false;
  for (const Cand& c : cands) {
    if (forced >= 0 ? (forced != c.variant) : c.only_forced) continue;
    if (nrt > c.rg * c.rt) continue;
    P->variant = c.variant; P->rg = c.rg; P->cg = c.cg; P->rt = c.rt; P->ct = c.ct;
    if (plan_variant(L, c.limit, P) == 0) { ok = true; break; }
  }
  if (!ok) return -1;
  const int C = P->C;
  assign_tiles(nrt, P->rg, P->rt, P->tile_id);
  // Work items: (candidate tile, chunk of hyper-samples), launched chunk-major, one CTA per SM at a time.  An item costs
  // ~31k cycles per sample + ~2.1k to start (measured at config 3), and the grid finishes when the last CTA does, so the
  // schedule is what list scheduling wants: long items first, short ones at the end.
  //   tiles >= SMs: halving chunks S/2, S/4, ..., 1, 1 of every tile (simulated end-of-grid loss at config 3: 0.8 % on
  //                 one GPU, 1.1 % for the 12 500-candidate shard of 8 GPUs; equal chunks of 8 / 4: 1.5 % / 5.6 %)
  //   tiles <  SMs: equal chunks, the size that minimises ceil(items / SMs) * (sc * 15 + 1)
  //   BODE_WAVES=w: round 1's rule (equal chunks, at least w waves of CTAs), kept as a tuning knob
  const int64_t ntiles = (Nd + C - 1) / C;
  int sc = L.S;
  P->n_tail = 0;
  P->tail_s0[0] = 0;
  if (tuning().waves > 0) {
    const int64_t target = (int64_t)tuning().waves * sm_count;
    while (sc > 4 && ntiles * ((L.S + sc - 1) / sc) < target) sc = (sc + 1) / 2;
    if (ntiles * ((L.S + sc - 1) / sc) < sm_count) {
      while (sc > 1 && ntiles * ((L.S + sc - 1) / sc) < sm_count) sc = (sc + 1) / 2;
    }
    P->n_main = (L.S + sc - 1) / sc;
  } else if (ntiles >= sm_count && L.S > 1) {
    int rem = L.S, pos = 0, nt = 0;
    while (rem > 0 && nt < kMaxTail) {
      const int c = (nt == kMaxTail - 1) ? rem : (rem + 1) / 2;
      P->tail_s0[nt++] = pos;
      pos += c;
      rem -= c;
    }
    P->tail_s0[nt] = pos;
    P->n_tail = nt;
    P->n_main = 0;
    sc = (L.S + 1) / 2;  // (reported: the largest item)
  } else {
    int64_t best_cost = -1;
    for (int c = 1; c <= L.S; c = (c < L.S && 2 * c > L.S) ? L.S : 2 * c) {
      const int64_t items = ntiles * ((L.S + c - 1) / c);
      const int64_t cost = ((items + sm_count - 1) / sm_count) * ((int64_t)c * 15 + 1);
      if (best_cost < 0 || cost <= best_cost) { best_cost = cost; sc = c; }  // ties: the larger chunk
      if (c == L.S) break;
    }
    P->n_main = (L.S + sc - 1) / sc;
  }
  P->sc = sc;
  P->ntiles = (int)ntiles;
  P->nitems = ntiles * (P->n_main + P->n_tail);
  return 0;
}

template <int RG, int CG, int RT, int CT>
static cudaError_t launch_one(const ScoreArgs& a, const ScorePlan& P, cudaStream_t st) {
  static thread_local int attr_dev = -1;
  static thread_local size_t attr_smem = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_dev || P.smem > attr_smem) {  // opt in to the large dynamic shared memory once per device / size
    e = cudaFuncSetAttribute(k_score<RG, CG, RT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
    attr_smem = P.smem;
  }
  k_score<RG, CG, RT, CT><<<(unsigned)P.nitems, 32 * RG * CG, P.smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_score(const ScoreArgs& a, const ScorePlan& P, cudaStream_t st) {
  if (P.variant == 0) return launch_one<4, 4, 8, 2>(a, P, st);
  if (P.variant == 3) return launch_one<4, 2, 8, 2>(a, P, st);
  if (P.variant == 4) return launch_one<4, 2, 8, 4>(a, P, st);
  if (P.variant == 6) return launch_one<4, 2, 16, 2>(a, P, st);
  if (P.variant == 7) return launch_one<8, 1, 16, 2>(a, P, st);
  if (P.variant == 5) return launch_one<16, 1, 8, 1>(a, P, st);
  if (P.variant == 1) return launch_one<4, 4, 16, 1>(a, P, st);
  return launch_one<8, 2, 16, 1>(a, P, st);
}

cudaError_t launch_ekld_reduce(const void* ws, const Layout& L, const double* Xd, const double* xi_ovr, double* sx,
                               const double* wkplane, int S, int64_t Nd, int64_t idx_offset, int form, int keep_logk,
                               double* score, double* var, double* part_val, int64_t* part_idx, unsigned int* done_counter,
                               double* best, int64_t* best_idx, cudaStream_t st) {
  const int cpl = Nd >= 64 * 148 * 2 ? 2 : 1;  // small designs: one candidate per lane, twice the blocks
  const int nb = (int)((Nd + 32 * cpl - 1) / (32 * cpl));
  const size_t smem = sizeof(double) * kErWarps * (size_t)L.d * kXcStride;  // 25.6 KB at d = 8, 51.2 KB at d = 16
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_dev) {  // d = 16 needs more than the 48 KB default: opt in once per device
    const int max_smem = (int)(sizeof(double) * kErWarps * 16 * kXcStride);
    e = cudaFuncSetAttribute(k_ekld_reduce<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_ekld_reduce<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
  }
  if (cpl == 2)
    k_ekld_reduce<2><<<nb, 32 * kErWarps, smem, st>>>(static_cast<const char*>(ws), L, Xd, xi_ovr, sx, wkplane, S, Nd, idx_offset, form,
                                                      tuning().xik_series, keep_logk || var != nullptr, score, var, part_val, part_idx,
                                                      done_counter, best, best_idx);
  else
    k_ekld_reduce<1><<<nb, 32 * kErWarps, smem, st>>>(static_cast<const char*>(ws), L, Xd, xi_ovr, sx, wkplane, S, Nd, idx_offset, form,
                                                      tuning().xik_series, keep_logk || var != nullptr, score, var, part_val, part_idx,
                                                      done_counter, best, best_idx);
  return cudaGetLastError();
}

cudaError_t launch_reduce(const double* logk, int S, int64_t Nd, int64_t idx_offset, double* score, double* var,
                          double* part_val, int64_t* part_idx, unsigned int* done_counter, double* best, int64_t* best_idx,
                          cudaStream_t st) {
  const int nb = (int)((Nd + 255) / 256);
  k_reduce<<<nb, 256, 0, st>>>(logk, S, Nd, idx_offset, score, var, part_val, part_idx, done_counter, best, best_idx);
  return cudaGetLastError();
}

cudaError_t launch_exp_inplace(double* v, size_t n, cudaStream_t st) {
  k_exp_inplace<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n);
  return cudaGetLastError();
}

}  // namespace bode

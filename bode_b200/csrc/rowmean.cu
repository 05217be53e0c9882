// K5 -- kernel row means over quadrature points (quadrature mode of the north star):
//     out[s][i] = (variance_s / sum_q w_q) * sum_q w_q exp(-sum_k ((P[i][k] - Xq[q][k]) / (sqrt2 ell_sk))^2)
// i.e. the quadrature estimate of  int k_s(P_i, x') dx'  that replaces the closed-form xik (reference _core.py:86-92)
// for the observations (e), the candidates (xi) and -- with P = Xq -- the double integral bk (:95-101).
// There is no reference call site (the reference integrates in closed form); parity is against the oracle's own
// quadrature variant on the same X_quad.
//
// Two points per thread, one hyper-sample per blockIdx.y; quadrature points are staged through shared memory in the
// scaled, centred coordinates of rbf.h (doubled, with -|v|^2 appended); each thread sums q = 0..Nq-1 into four interleaved
// accumulators (q mod 4) combined in a fixed order, so the result does not depend on the launch geometry.  FP64-pipe
// bound: d + 11 operations per (point, quadrature point) pair (GEMM-form distance + 9-operation table exp2).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "layout.h"
#include "rbf.h"

namespace bode {

constexpr int kRmThreads = 256;
constexpr int kRmChunk = 512;  // quadrature points per shared-memory chunk
constexpr int kRmPts = 2;      // points per thread (the staged quadrature rows are read once for both)

// z = -|u_i|^2 - |v_q|^2 + 2 u_i . v_q in the scaled, centred coordinates of rbf.h (u = kappa c (p - 1/2)), the same
// GEMM-form distance and 9-operation table exp2 as the scoring kernel's cross-covariance: d + 11 FP64 operations per
// (point, quadrature point) pair instead of the 2d + 12 of direct differences + the 11-operation exp.
template <int DP, bool WEIGHTED>
__device__ __forceinline__ void rowmean_body(const double* __restrict__ P, int64_t Np, const double* __restrict__ Xq,
                                             const double* __restrict__ wq, int Nq, int d, const double* cku,
                                             const double* tab, double* qs, double* ws_, double scale, double* __restrict__ out) {
  constexpr int DPAD = (DP + 2) & ~1;  // DP doubled coordinates + -|v|^2, padded to an even count (16-byte rows)
  const double* tab_lane = tab + (threadIdx.x & (kTabRep - 1));  // bank-staggered table copies: conflict-free lookups
  int64_t i[kRmPts];
  double y[kRmPts][DP], nyn[kRmPts], acc[kRmPts][4];
#pragma unroll
  for (int p = 0; p < kRmPts; p++) {
    i[p] = ((int64_t)blockIdx.x * kRmPts + p) * kRmThreads + threadIdx.x;
    nyn[p] = 0.0;
#pragma unroll
    for (int k = 0; k < DP; k++) {
      y[p][k] = (i[p] < Np) ? (P[i[p] * d + k] - 0.5) * cku[k] : 0.0;
      nyn[p] = fma(-y[p][k], y[p][k], nyn[p]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) acc[p][j] = 0.0;
  }
  for (int q0 = 0; q0 < Nq; q0 += kRmChunk) {
    const int nq = min(kRmChunk, Nq - q0);
    __syncthreads();
    // staged row of quadrature point q: [2 v_0 .. 2 v_(DP-1), -|v|^2]; rows beyond nq: far away (k = 0 exactly after the clamp)
    for (int q = threadIdx.x; q < kRmChunk; q += kRmThreads) {
      double nv = 0.0;
#pragma unroll
      for (int k = 0; k < DP; k++) {
        const double v = (q < nq) ? (Xq[(size_t)(q0 + q) * d + k] - 0.5) * cku[k] : 0.0;
        qs[(size_t)q * DPAD + k] = 2.0 * v;
        nv = fma(-v, v, nv);
      }
      qs[(size_t)q * DPAD + DP] = nv;
      if (WEIGHTED) ws_[q] = (q < nq) ? wq[q0 + q] : 0.0;
    }
    __syncthreads();
    const int nq4 = (nq + 3) & ~3;  // padded rows are masked below (unweighted) or carry weight 0 (weighted)
    for (int q = 0; q < nq4; q += 4) {
#pragma unroll
      for (int p = 0; p < kRmPts; p++) {
        double z[4], kv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double* row = qs + (size_t)(q + j) * DPAD;
          double t = row[DP] + nyn[p];
#pragma unroll
          for (int k = 0; k < DP; k++) t = fma(y[p][k], row[k], t);
          z[j] = fmin(t, 0.0);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) kv[j] = rbf_exp64<kTabRep>(z[j], tab_lane);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (WEIGHTED) acc[p][j] = fma(ws_[q + j], kv[j], acc[p][j]);
          else acc[p][j] += (q + j < nq) ? kv[j] : 0.0;
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < kRmPts; p++)
    if (i[p] < Np) out[i[p]] = ((acc[p][0] + acc[p][1]) + (acc[p][2] + acc[p][3])) * scale;
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(kRmThreads) k_rowmean(const double* __restrict__ hyp, int ndim, int d,
                                                        const double* __restrict__ P, int64_t Np,
                                                        const double* __restrict__ Xq, const double* __restrict__ wq,
                                                        int Nq, const double* __restrict__ wsum_ptr, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  double* qs = sm;                               // [kRmChunk][dpad]
  double* ws_ = qs + kRmChunk * ((d + 2) & ~1);  // [kRmChunk]
  double* tab = ws_ + kRmChunk;                  // [64][16] 2^(j/64), each entry replicated 16 times (layout.h): WITHOUT the
                                                 // variance (scale applied at the end)
  double* cku = tab + kTab64 * kTabRep;          // [16] kappa / (sqrt2 ell_k)
  const int s = blockIdx.y;
  const double* h = hyp + (size_t)s * ndim;
  for (int idx = threadIdx.x; idx < kTab64 * kTabRep; idx += kRmThreads) tab[idx] = exp2((double)(idx / kTabRep) / 64.0);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kCkPad) {
    const int k = threadIdx.x - 64;
    cku[k] = (k < d) ? kKappa / (1.4142135623730951 * h[1 + k]) : 0.0;
  }
  __syncthreads();
  const double scale = h[0] / (wsum_ptr ? wsum_ptr[0] : (double)Nq);
  double* o = out + (size_t)s * Np;
  switch (d) {
#define BODE_CASE(D)                                                                            \
  case D:                                                                                       \
    rowmean_body<D, WEIGHTED>(P, Np, Xq, wq, Nq, d, cku, tab, qs, ws_, scale, o);               \
    break;
    BODE_CASE(1) BODE_CASE(2) BODE_CASE(3) BODE_CASE(4) BODE_CASE(5) BODE_CASE(6) BODE_CASE(7) BODE_CASE(8)
    BODE_CASE(9) BODE_CASE(10) BODE_CASE(11) BODE_CASE(12) BODE_CASE(13) BODE_CASE(14) BODE_CASE(15) BODE_CASE(16)
#undef BODE_CASE
    default:
      break;
  }
}

// sigma0[s] = sum_q w_q rowmean[s][q] / wsum   (fixed order: one thread per sample, sequential in q)
__global__ void k_weighted_mean(const double* __restrict__ rm, const double* __restrict__ wq, int Nq,
                                const double* __restrict__ wsum_ptr, int S, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double acc = 0.0;
  for (int q = 0; q < Nq; q++) acc = fma(wq ? wq[q] : 1.0, rm[(size_t)s * Nq + q], acc);
  out[s] = acc / (wsum_ptr ? wsum_ptr[0] : (double)Nq);
}

// sum of the quadrature weights, fixed order (thread t sums q == t mod 256, then an ordered pass)
__global__ void __launch_bounds__(256) k_sum_weights(const double* __restrict__ wq, int Nq, double* __restrict__ out) {
  __shared__ double part[256];
  double a = 0.0;
  for (int q = threadIdx.x; q < Nq; q += 256) a += wq[q];
  part[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 256; i++) t += part[i];
    out[0] = t;
  }
}

cudaError_t launch_rowmean(const double* hyp, int S, int ndim, int d, const double* P, int64_t Np, const double* Xq,
                           const double* wq, int Nq, const double* wsum_ptr, double* out, cudaStream_t st) {
  const int dpad = (d + 2) & ~1;
  const size_t smem = sizeof(double) * ((size_t)kRmChunk * dpad + kRmChunk + kTab64 * kTabRep + kCkPad);
  const dim3 grid((unsigned)((Np + (int64_t)kRmThreads * kRmPts - 1) / ((int64_t)kRmThreads * kRmPts)), (unsigned)S);
  // (the largest configuration, d = 16, needs 78 KB: opt in once per device)
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_dev) {
    const int max_smem = (int)(sizeof(double) * ((size_t)kRmChunk * 18 + kRmChunk + kTab64 * kTabRep + kCkPad));
    e = cudaFuncSetAttribute(k_rowmean<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_rowmean<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
  }
  if (wq) k_rowmean<true><<<grid, kRmThreads, smem, st>>>(hyp, ndim, d, P, Np, Xq, wq, Nq, wsum_ptr, out);
  else k_rowmean<false><<<grid, kRmThreads, smem, st>>>(hyp, ndim, d, P, Np, Xq, wq, Nq, nullptr, out);
  return cudaGetLastError();
}

cudaError_t launch_weighted_mean(const double* rm, const double* wq, int Nq, const double* wsum_ptr, int S, double* out,
                                 cudaStream_t st) {
  k_weighted_mean<<<(S + 63) / 64, 64, 0, st>>>(rm, wq, Nq, wq ? wsum_ptr : nullptr, S, out);
  return cudaGetLastError();
}

cudaError_t launch_sum_weights(const double* wq, int Nq, double* out, cudaStream_t st) {
  k_sum_weights<<<1, 256, 0, st>>>(wq, Nq, out);
  return cudaGetLastError();
}

}  // namespace bode

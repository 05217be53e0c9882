// K5 -- kernel row means over quadrature points (quadrature mode of the north star):
//     out[s][i] = (variance_s / sum_q w_q) * sum_q w_q exp(-sum_k ((P[i][k] - Xq[q][k]) / (sqrt2 ell_sk))^2)
// i.e. the quadrature estimate of  int k_s(P_i, x') dx'  that replaces the closed-form xik (reference _core.py:86-92)
// for the observations (e), the candidates (xi) and -- with P = Xq -- the double integral bk (:95-101).
// There is no reference call site (the reference integrates in closed form); parity is against the oracle's own
// quadrature variant on the same X_quad.
//
// One thread per point, one hyper-sample per blockIdx.y; quadrature points are staged through shared memory already
// scaled by 1/(sqrt2 ell); each thread sums q = 0..Nq-1 into four interleaved accumulators (q mod 4) combined in a
// fixed order, so the result does not depend on the launch geometry.  FP64-pipe bound: 2d + 12 operations per
// (point, quadrature point) pair.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "layout.h"
#include "rbf.h"

namespace bode {

constexpr int kRmThreads = 256;
constexpr int kRmChunk = 512;  // quadrature points per shared-memory chunk

template <int DP, bool WEIGHTED>
__device__ __forceinline__ void rowmean_body(const double* __restrict__ P, int64_t Np, const double* __restrict__ Xq,
                                             const double* __restrict__ wq, int Nq, int d, const double* ck,
                                             const double* tab, double* qs, double* ws_, double scale, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kRmThreads + threadIdx.x;
  double y[DP];
#pragma unroll
  for (int k = 0; k < DP; k++) y[k] = (i < Np) ? P[i * d + k] * ck[k] : 0.0;
  constexpr int DPAD = (DP + 1) & ~1;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int q0 = 0; q0 < Nq; q0 += kRmChunk) {
    const int nq = min(kRmChunk, Nq - q0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kRmChunk * DPAD; idx += kRmThreads) {
      const int q = idx / DPAD, k = idx - q * DPAD;
      qs[idx] = (q < nq && k < d) ? Xq[(size_t)(q0 + q) * d + k] * ck[k] : 0.0;
    }
    if (WEIGHTED)
      for (int q = threadIdx.x; q < kRmChunk; q += kRmThreads) ws_[q] = (q < nq) ? wq[q0 + q] : 0.0;
    __syncthreads();
    const int nq4 = (nq + 3) & ~3;  // padded rows contribute exp(-|y|^2) * 0 (weighted) or are masked below
    for (int q = 0; q < nq4; q += 4) {
      double kv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) kv[j] = rbf_exp(-sqdist<DP>(qs + (size_t)(q + j) * DPAD, y), tab);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (WEIGHTED) acc[j] = fma(ws_[q + j], kv[j], acc[j]);
        else acc[j] += (q + j < nq) ? kv[j] : 0.0;
      }
    }
  }
  if (i < Np) out[i] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) * scale;
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(kRmThreads) k_rowmean(const double* __restrict__ hyp, int ndim, int d,
                                                        const double* __restrict__ P, int64_t Np,
                                                        const double* __restrict__ Xq, const double* __restrict__ wq,
                                                        int Nq, const double* __restrict__ wsum_ptr, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  double* qs = sm;                               // [kRmChunk][dpad]
  double* ws_ = qs + kRmChunk * ((d + 1) & ~1);  // [kRmChunk]
  double* tab = ws_ + kRmChunk;                  // [32] exp table WITHOUT the variance (scale applied at the end)
  double* ck = tab + kTabPad;                    // [16]
  const int s = blockIdx.y;
  const double* h = hyp + (size_t)s * ndim;
  if (threadIdx.x < kTabPad) tab[threadIdx.x] = exp2((double)threadIdx.x / 32.0);
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kCkPad) {
    const int k = threadIdx.x - 32;
    ck[k] = (k < d) ? 1.0 / (1.4142135623730951 * h[1 + k]) : 0.0;
  }
  __syncthreads();
  const double scale = h[0] / (wsum_ptr ? wsum_ptr[0] : (double)Nq);
  double* o = out + (size_t)s * Np;
  switch (d) {
#define BODE_CASE(D)                                                                            \
  case D:                                                                                       \
    rowmean_body<D, WEIGHTED>(P, Np, Xq, wq, Nq, d, ck, tab, qs, ws_, scale, o);                \
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
  const int dpad = (d + 1) & ~1;
  const size_t smem = sizeof(double) * ((size_t)kRmChunk * dpad + kRmChunk + kTabPad + kCkPad);
  const dim3 grid((unsigned)((Np + kRmThreads - 1) / kRmThreads), (unsigned)S);
  // (the largest configuration, d = 16, needs 72 KB: opt in once per device)
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_dev) {
    const int max_smem = (int)(sizeof(double) * ((size_t)kRmChunk * 16 + kRmChunk + kTabPad + kCkPad));
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

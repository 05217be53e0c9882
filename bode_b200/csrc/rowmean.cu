// K5 -- kernel row means over quadrature points (quadrature mode of the north star):
//     out[s][i] = (variance_s / sum_q w_q) * sum_q w_q exp(-sum_k ((P[i][k] - Xq[q][k]) / (sqrt2 ell_sk))^2)
// i.e. the quadrature estimate of  int k_s(P_i, x') dx'  that replaces the closed-form xik (reference _core.py:86-92)
// for the observations (e), the candidates (xi) and -- with P = Xq -- the double integral bk (:95-101).
// There is no reference call site (the reference integrates in closed form); parity is against the oracle's own
// quadrature variant on the same X_quad.
//
// Two kernels, both with a launch-geometry-independent summation order:
//   k_rowmean_mma (default)  8 x 8 (point, quadrature point) tiles on the FP64 tensor pipe: GEMM-form squared distance by
//                            ceil(d/4) DMMA.8x8x4 per tile, the 9-operation table exp2 on the accumulators, weighted column
//                            sums.  6.1e11 pairs/s at d = 10, 7.1e11 at d = 8 (B200; round 1's kernel: 4.4e11).
//   k_rowmean                thread-per-point form (BODE_ROWMEAN_VARIANT=0): two points per thread, quadrature rows staged
//                            through shared memory, d + 11 FP64 operations per pair (5.4e11 pairs/s at d = 10).
// FP64-pipe accounting at d = 10, per 64 pairs and scheduler: 3 DMMA x 16 + 22 DFMA/DADD x 2 = 92 cycles (8.1e11 pairs/s at
// 100 %); the exponential alone is 36 of them, so 2x round 1 (8.8e11) is above what this pipe can deliver.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "layout.h"
#include "rbf.h"
#include "tuning.h"

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

// ---- tensor-pipe variant: 8 x 8 (point, quadrature point) tiles, z = -|u|^2 - |v|^2 + 2 u . v by DK = ceil(d/4) DMMA.8x8x4 per
// tile (A fragments: the warp's 32 points, held in registers; B fragments: the staged quadrature tile), then the 9-operation
// table exp2 on the accumulators and a weighted column sum.  FP64-pipe cost per 64 pairs: 16 DK cycles of DMMA + 44 cycles
// of DFMA/DADD, against 2 (d + 11) = 42 two-cycle operations of the thread-per-point form -- and a third of its instructions.
// Fixed summation order: lane (r, c) adds the columns 2c, 2c + 1 of the quadrature tiles in ascending order, the four lanes
// of a row are combined by a fixed tree: the result does not depend on the launch geometry.
constexpr int kRmmWarps = 8;
constexpr int kRmmRT = 4;       // row tiles (of 8 points) per warp
constexpr int kRmmChunk = 256;  // quadrature points per shared-memory chunk (32 tiles; 512 measured no faster)

__device__ __forceinline__ void rmm_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int DK>
__global__ void __launch_bounds__(32 * kRmmWarps, 2) k_rowmean_mma(const double* __restrict__ hyp, int ndim, int d,
                                                                    const double* __restrict__ P, int64_t Np,
                                                                    const double* __restrict__ Xq, const double* __restrict__ wq,
                                                                    int Nq, const double* __restrict__ wsum_ptr,
                                                                    double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  constexpr int kTileDoubles = DK * 32 + 16;     // B fragments of the DK k-steps, then {-|v|^2}[8], {w}[8]
  double* qt = sm;                                // [kRmmChunk / 8][kTileDoubles]
  double* tab = qt + (kRmmChunk / 8) * kTileDoubles;  // [64][16] 2^(j/64), bank-staggered copies (layout.h)
  double* cku = tab + kTab64 * kTabRep;           // [16]
  const int s = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = lane >> 2, c = lane & 3;
  const double* h = hyp + (size_t)s * ndim;
  for (int idx = tid; idx < kTab64 * kTabRep; idx += 32 * kRmmWarps) tab[idx] = exp2((double)(idx / kTabRep) / 64.0);
  if (tid < kCkPad) cku[tid] = (tid < d) ? kKappa / (1.4142135623730951 * h[1 + tid]) : 0.0;
  __syncthreads();
  const unsigned tab_lane = (unsigned)__cvta_generic_to_shared(tab + (lane & (kTabRep - 1)));
  // this warp's points: A fragments (row r of tile t, dimension 4i + c) and -|u|^2 of row r
  const int64_t p0 = ((int64_t)blockIdx.x * kRmmWarps + warp) * (8 * kRmmRT);
  double af[kRmmRT][DK], nun[kRmmRT], acc[kRmmRT][2];
#pragma unroll
  for (int t = 0; t < kRmmRT; t++) {
    const int64_t pi = p0 + 8 * t + r;
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < DK; i++) {
      const int k = 4 * i + c;
      const double u = (pi < Np && k < d) ? (P[pi * d + k] - 0.5) * cku[k] : 0.0;
      af[t][i] = u;
      part = fma(-u, u, part);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);  // the four lanes of a row hold its dimensions c, c + 4, ...
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    nun[t] = part;
    acc[t][0] = 0.0;
    acc[t][1] = 0.0;
  }
  for (int q0 = 0; q0 < Nq; q0 += kRmmChunk) {
    __syncthreads();
    // stage the chunk: B fragment (k = l & 3, n = l >> 2) of tile j, k-step i = 2 v[q0 + 8j + n][4i + k]; rows beyond Nq: weight 0
    for (int idx = tid; idx < (kRmmChunk / 8) * DK * 32; idx += 32 * kRmmWarps) {
      const int j = idx / (DK * 32), rem = idx - j * (DK * 32), i = rem >> 5, l = rem & 31;
      const int q = q0 + 8 * j + (l >> 2), k = 4 * i + (l & 3);
      qt[j * kTileDoubles + i * 32 + l] = (q < Nq && k < d) ? 2.0 * ((Xq[(size_t)q * d + k] - 0.5) * cku[k]) : 0.0;
    }
    for (int idx = tid; idx < kRmmChunk; idx += 32 * kRmmWarps) {
      const int q = q0 + idx;
      double nv = 0.0;
      if (q < Nq)
        for (int k = 0; k < d; k++) {
          const double v = (Xq[(size_t)q * d + k] - 0.5) * cku[k];
          nv = fma(-v, v, nv);
        }
      double* tl = qt + (idx >> 3) * kTileDoubles + DK * 32;
      tl[idx & 7] = nv;
      tl[8 + (idx & 7)] = (q < Nq) ? (wq ? wq[q] : 1.0) : 0.0;
    }
    __syncthreads();
    const int ntile = (min(kRmmChunk, Nq - q0) + 7) >> 3;
    for (int j = 0; j < ntile; j++) {
      const double* tl = qt + j * kTileDoubles;
      double b[DK];
#pragma unroll
      for (int i = 0; i < DK; i++) b[i] = tl[i * 32 + lane];
      const double2 nv = *reinterpret_cast<const double2*>(tl + DK * 32 + 2 * c);
      const double2 wv = *reinterpret_cast<const double2*>(tl + DK * 32 + 8 + 2 * c);
      double z[2 * kRmmRT], kv[2 * kRmmRT];
#pragma unroll
      for (int t = 0; t < kRmmRT; t++) {
        z[2 * t] = nun[t] + nv.x;
        z[2 * t + 1] = nun[t] + nv.y;
      }
#pragma unroll
      for (int i = 0; i < DK; i++)
#pragma unroll
        for (int t = 0; t < kRmmRT; t++) rmm_dmma(z[2 * t], z[2 * t + 1], af[t][i], b[i]);
      rbf_exp64_n<2 * kRmmRT, kTabRep>(z, tab_lane, kv);
#pragma unroll
      for (int t = 0; t < kRmmRT; t++) {
        acc[t][0] = fma(wv.x, kv[2 * t], acc[t][0]);
        acc[t][1] = fma(wv.y, kv[2 * t + 1], acc[t][1]);
      }
    }
  }
  const double scale = h[0] / (wsum_ptr ? wsum_ptr[0] : (double)Nq);
#pragma unroll
  for (int t = 0; t < kRmmRT; t++) {
    double v = acc[t][0] + acc[t][1];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    const int64_t pi = p0 + 8 * t + r;
    if (c == 0 && pi < Np) out[(size_t)s * Np + pi] = v * scale;
  }
}

template <int DK>
static cudaError_t launch_rowmean_mma_dk(const double* hyp, int S, int ndim, int d, const double* P, int64_t Np, const double* Xq,
                                         const double* wq, int Nq, const double* wsum_ptr, double* out, cudaStream_t st) {
  const size_t smem = sizeof(double) * ((size_t)(kRmmChunk / 8) * (DK * 32 + 16) + kTab64 * kTabRep + kCkPad);
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_dev) {
    e = cudaFuncSetAttribute(k_rowmean_mma<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
  }
  const int64_t per_block = (int64_t)kRmmWarps * 8 * kRmmRT;
  const dim3 grid((unsigned)((Np + per_block - 1) / per_block), (unsigned)S);
  k_rowmean_mma<DK><<<grid, 32 * kRmmWarps, smem, st>>>(hyp, ndim, d, P, Np, Xq, wq, Nq, wsum_ptr, out);
  return cudaGetLastError();
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
  if (tuning().rowmean_variant != 0) {  // tensor-pipe variant (default); BODE_ROWMEAN_VARIANT=0: thread-per-point form below
    switch ((d + 3) / 4) {
      case 1: return launch_rowmean_mma_dk<1>(hyp, S, ndim, d, P, Np, Xq, wq, Nq, wsum_ptr, out, st);
      case 2: return launch_rowmean_mma_dk<2>(hyp, S, ndim, d, P, Np, Xq, wq, Nq, wsum_ptr, out, st);
      case 3: return launch_rowmean_mma_dk<3>(hyp, S, ndim, d, P, Np, Xq, wq, Nq, wsum_ptr, out, st);
      default: return launch_rowmean_mma_dk<4>(hyp, S, ndim, d, P, Np, Xq, wq, Nq, wsum_ptr, out, st);
    }
  }
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

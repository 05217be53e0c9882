// FP64 pipe microbenchmarks for B200 (sm_100a): DFMA, DMMA.8x8x4, mixed issue, libm
// transcendentals and an smem-fed DMMA loop.  Design evidence for bode_b200's scoring kernel
// (DESIGN.md section "Why DMMA"); not part of the product path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_microbench tools/fp64_microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double x, double y, long long* cyc) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double x, double y, long long* cyc) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  double a = x + threadIdx.x * 1e-9, b = y;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// same warp interleaves NM DMMAs with NF DFMAs per iteration
template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double x, double y, long long* cyc) {
  double c0[NM], c1[NM], f[NF];
#pragma unroll
  for (int i = 0; i < NM; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = threadIdx.x * 1e-3 + i;
  double a = x + threadIdx.x * 1e-9, b = y;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < (NM > NF ? NM : NF); i++) {
      if (i < NM) dmma884(c0[i], c1[i], a, b);
      if (i < NF) f[i] = fma(f[i], x, y);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// warp-specialised: even warps DMMA, odd warps DFMA
__global__ void __launch_bounds__(256) k_split(double* out, int iters, double x, double y, long long* cyc) {
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  double a = x + threadIdx.x * 1e-9, b = y;
  long long t0 = clock64();
  if ((threadIdx.x >> 5) & 1) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) { c0[i] = fma(c0[i], x, y); c1[i] = fma(c1[i], x, y); }
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma884(c0[i], c1[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(256) k_libm(double* out, int iters, double x, long long* cyc) {
  double v[4];
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = -(threadIdx.x * 0.01 + i * 0.37 + x);
  double s = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double r;
      if (OP == 0) r = exp(v[i]);
      else if (OP == 1) r = erf(v[i] * 0.1);
      else if (OP == 2) r = log(-v[i] + 1.5);
      else if (OP == 3) r = sqrt(-v[i] + 1.5);
      else r = 1.0 / (-v[i] + 1.5);
      s += r;
      v[i] -= 1e-3;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}


// exp(x) for x <= 0: Cody-Waite reduction + degree-13 Horner, 2^k applied by exponent add.
__device__ __forceinline__ double exp_neg(double x) {
  x = fmax(x, -708.0);
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  double t = fma(x, 1.4426950408889634, SHIFT);
  double k = t - SHIFT;
  double r = fma(k, -6.93147180369123816490e-01, x);
  r = fma(k, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;           // 1/13!
  p = fma(p, r, 2.08767569878681e-09);          // 1/12!
  p = fma(p, r, 2.505210838544172e-08);         // 1/11!
  p = fma(p, r, 2.755731922398589e-07);         // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
  p = fma(p, r, 2.48015873015873e-05);          // 1/8!
  p = fma(p, r, 1.984126984126984e-04);         // 1/7!
  p = fma(p, r, 1.388888888888889e-03);         // 1/6!
  p = fma(p, r, 8.333333333333333e-03);         // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int hi = __double2hiint(p) + (__double2loint(t) << 20);
  return __hiloint2double(hi, __double2loint(p));
}
__global__ void __launch_bounds__(256) k_exp_custom(double* out, int iters, double x, long long* cyc) {
  double v[4];
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = -(threadIdx.x * 0.01 + i * 0.37 + x);
  double s = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) { s += exp_neg(v[i]); v[i] -= 1e-3; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_exp_check(double* err, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = -720.0 * (double)i / n;
  double a = exp_neg(x), b = exp(x);
  err[i] = (b > 1e-300) ? fabs(a - b) / b : 0.0;
}

// smem-fed DMMA: warp tile RT row-tiles x CT col-tiles; per k4-step loads RT A-frags + CT B-frags (LDS.64)
template <int RT, int CT, int NT = 256>
__global__ void __launch_bounds__(NT) k_dmma_smem(double* out, int iters, long long* cyc) {
  extern __shared__ double sm[];
  const int KG = 16;  // k4 groups resident
  double* sA = sm;                          // [KG][RT*8 rows][4]
  constexpr int NCG = (NT == 256) ? 2 : 4;
  double* sB = sm + KG * RT * 32 * 4;       // [KG][CT*8 cols * NCG col-groups][4]
  for (int i = threadIdx.x; i < KG * RT * 32 * 4 + KG * CT * 8 * NCG * 4; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = (NT == 256) ? (warp & 3) : (warp >> 2), cg = (NT == 256) ? (warp >> 2) : (warp & 3);
  double c0[RT][CT], c1[RT][CT];
#pragma unroll
  for (int r = 0; r < RT; r++)
#pragma unroll
    for (int c = 0; c < CT; c++) { c0[r][c] = 0; c1[r][c] = 0; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    for (int g = 0; g < KG; g++) {
      double a[RT], b[CT];
      const double* pa = sA + (g * RT * 32 + rg * RT * 8) * 4 + lane;  // rows (lane/4), k (lane%4) contiguous 256B per tile
      const double* pb = sB + (g * CT * 8 * NCG + cg * CT * 8) * 4 + lane;
#pragma unroll
      for (int r = 0; r < RT; r++) a[r] = pa[r * 32];
#pragma unroll
      for (int c = 0; c < CT; c++) b[c] = pb[c * 32];
#pragma unroll
      for (int r = 0; r < RT; r++)
#pragma unroll
        for (int c = 0; c < CT; c++) dmma884(c0[r][c], c1[r][c], a[r], b[c]);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int r = 0; r < RT; r++)
#pragma unroll
    for (int c = 0; c < CT; c++) s += c0[r][c] + c1[r][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  void start() { CK(cudaEventRecord(a)); }
  float stop() { CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};


// DADD / DMUL issue rate (are they full-rate like DFMA?) and an int-heavy mix like the table exp2 of phase A
template <int OP, int NACC>
__global__ void __launch_bounds__(256) k_dop(double* out, int iters, double x, double y, long long* cyc) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i + 1.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (OP == 0) acc[i] = __dadd_rn(acc[i], y);
      else if (OP == 1) acc[i] = __dmul_rn(acc[i], x);
      else { acc[i] = __dadd_rn(acc[i], y); acc[i] = __fma_rn(acc[i], x, y); }  // alternating DADD / DFMA
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// NF DFMA + NI integer operations per accumulator per iteration (does integer work steal FP64 issue slots?)
template <int NACC, int NI>
__global__ void __launch_bounds__(256) k_dfma_int(double* out, int iters, double x, double y, long long* cyc) {
  double acc[NACC];
  int v[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { acc[i] = threadIdx.x * 1e-3 + i; v[i] = threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      acc[i] = fma(acc[i], x, y);
#pragma unroll
      for (int j = 0; j < NI; j++) v[i] = (v[i] ^ (v[i] >> 3)) + it;   // 2-3 integer instructions
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i] + v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", p.name, sms, p.major, p.minor);
  double* out; long long* cyc; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024)); CK(cudaMalloc(&cyc, 8));
  Timer T;
  long long hc;
#define RUN(name, flops_per_thread_iter, iters, blocks_per_sm, threads, launch)                          \
  do {                                                                                                    \
    int grid = sms * (blocks_per_sm);                                                                     \
    for (int w = 0; w < 2; w++) { launch; }                                                               \
    CK(cudaDeviceSynchronize());                                                                          \
    float best = 1e30f;                                                                                   \
    for (int r = 0; r < 5; r++) { T.start(); launch; float ms = T.stop(); if (ms < best) best = ms; }     \
    CK(cudaGetLastError());                                                                               \
    CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));                                                  \
    double total = (double)(flops_per_thread_iter) * (iters) * grid * (threads);                          \
    printf("{\"test\": \"%s\", \"ms\": %.4f, \"tflops_or_gops\": %.3f, \"per_sm_per_clk\": %.2f, \"eff_mhz\": %.0f}\n", \
           name, best, total / best * 1e-9, total / 2.0 / sms / (double)hc, (double)hc / best * 1e-3);    \
    fflush(stdout);                                                                                       \
  } while (0)

  const int IT = 20000;
  // flops per thread per iter: DFMA: 2*NACC ; DMMA 8x8x4: 256 FMA / 32 lanes = 8 FMA = 16 flop per lane per instr
  RUN("dfma_acc8_occ4x256", 2 * 8, IT, 4, 256, (k_dfma<8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_acc16_occ2x256", 2 * 16, IT, 2, 256, (k_dfma<16><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_acc4_occ8x256", 2 * 4, IT, 8, 256, (k_dfma<4><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dadd_acc8_occ4x256", 1 * 8, IT, 4, 256, (k_dop<0, 8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dmul_acc8_occ4x256", 1 * 8, IT, 4, 256, (k_dop<1, 8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dadd_dfma_alt_acc8_occ4x256", 3 * 8, IT, 4, 256, (k_dop<2, 8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dadd_acc8_occ1x256", 1 * 8, IT, 1, 256, (k_dop<0, 8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_acc8_occ1x256", 2 * 8, IT, 1, 256, (k_dfma<8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_int1_acc8_occ1x256", 2 * 8, IT, 1, 256, (k_dfma_int<8, 1><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_int2_acc8_occ1x256", 2 * 8, IT, 1, 256, (k_dfma_int<8, 2><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_int1_acc8_occ2x256", 2 * 8, IT, 2, 256, (k_dfma_int<8, 1><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_int2_acc8_occ2x256", 2 * 8, IT, 2, 256, (k_dfma_int<8, 2><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dmma_acc8_occ1x256", 16 * 8, IT, 1, 256, (k_dmma<8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dmma_acc8_occ2x256", 16 * 8, IT, 2, 256, (k_dmma<8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dmma_acc16_occ1x256", 16 * 16, IT, 1, 256, (k_dmma<16><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dmma_acc4_occ1x128", 16 * 4, IT, 1, 128, (k_dmma<4><<<grid, 128>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dmma_acc1_occ1x128_latency", 16 * 1, IT, 1, 128, (k_dmma<1><<<grid, 128>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("dfma_acc1_occ1x128_latency", 2 * 1, IT, 1, 128, (k_dfma<1><<<grid, 128>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("mix_8dmma_8dfma_occ1x256", 16 * 8 + 2 * 8, IT, 1, 256, (k_mix<8, 8><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("mix_8dmma_16dfma_occ1x256", 16 * 8 + 2 * 16, IT, 1, 256, (k_mix<8, 16><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("mix_8dmma_32dfma_occ1x256", 16 * 8 + 2 * 32, IT, 1, 256, (k_mix<8, 32><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("mix_4dmma_32dfma_occ1x256", 16 * 4 + 2 * 32, IT, 1, 256, (k_mix<4, 32><<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  // split: half the warps DMMA (16*8 per iter), half DFMA (2*16 per iter): average per thread = (128+32)/2
  RUN("split_warps_dmma_dfma_occ1x256", (16 * 8 + 2 * 16) / 2, IT, 1, 256, (k_split<<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  RUN("split_warps_dmma_dfma_occ2x256", (16 * 8 + 2 * 16) / 2, IT, 2, 256, (k_split<<<grid, 256>>>(out, IT, 1.0000001, 1e-9, cyc)));
  const int IL = 2000;
  // libm: count 1 "op" per call -> G calls/s  (tflops_or_gops column = 1e12 calls/s ; per_sm_per_clk = calls/2)
  RUN("libm_exp_occ4x256", 4, IL, 4, 256, (k_libm<0><<<grid, 256>>>(out, IL, 0.5, cyc)));
  RUN("libm_erf_occ4x256", 4, IL, 4, 256, (k_libm<1><<<grid, 256>>>(out, IL, 0.5, cyc)));
  RUN("libm_log_occ4x256", 4, IL, 4, 256, (k_libm<2><<<grid, 256>>>(out, IL, 0.5, cyc)));
  RUN("libm_sqrt_occ4x256", 4, IL, 4, 256, (k_libm<3><<<grid, 256>>>(out, IL, 0.5, cyc)));
  RUN("libm_div_occ4x256", 4, IL, 4, 256, (k_libm<4><<<grid, 256>>>(out, IL, 0.5, cyc)));
  RUN("custom_exp_occ4x256", 4, IL, 4, 256, (k_exp_custom<<<grid, 256>>>(out, IL, 0.5, cyc)));
  {
    int n = 1 << 20; double* err; CK(cudaMalloc(&err, n * 8)); k_exp_check<<<n / 256, 256>>>(err, n);
    double* h = (double*)malloc(n * 8); CK(cudaMemcpy(h, err, n * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < n; i++) if (h[i] > mx) mx = h[i];
    printf("{\"test\": \"custom_exp_max_rel_err_vs_libm\", \"value\": %.3e}\n", mx);
  }
  RUN("libm_exp_occ1x256", 4, IL, 1, 256, (k_libm<0><<<grid, 256>>>(out, IL, 0.5, cyc)));
  // smem-fed DMMA: flops per thread per iter = KG(16) * RT*CT * 16
  {
    const int ISM = 400;
    size_t sh74 = (16 * 7 * 32 * 4 + 16 * 4 * 16 * 4) * 8;
    CK(cudaFuncSetAttribute(k_dmma_smem<7, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh74));
    RUN("dmma_smem_rt7_ct4_occ1x256", 16 * 7 * 4 * 16, ISM, 1, 256, (k_dmma_smem<7, 4><<<grid, 256, sh74>>>(out, ISM, cyc)));
    size_t sh44 = (16 * 4 * 32 * 4 + 16 * 4 * 16 * 4) * 8;
    CK(cudaFuncSetAttribute(k_dmma_smem<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh44));
    RUN("dmma_smem_rt4_ct4_occ1x256", 16 * 4 * 4 * 16, ISM, 1, 256, (k_dmma_smem<4, 4><<<grid, 256, sh44>>>(out, ISM, cyc)));
    size_t sh24 = (16 * 2 * 32 * 4 + 16 * 4 * 16 * 4) * 8;
    CK(cudaFuncSetAttribute(k_dmma_smem<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh24));
    RUN("dmma_smem_rt2_ct4_occ1x256", 16 * 2 * 4 * 16, ISM, 1, 256, (k_dmma_smem<2, 4><<<grid, 256, sh24>>>(out, ISM, cyc)));
    size_t sh14 = (16 * 1 * 32 * 4 + 16 * 4 * 16 * 4) * 8;
    CK(cudaFuncSetAttribute(k_dmma_smem<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh14));
    RUN("dmma_smem_rt1_ct4_occ1x256", 16 * 1 * 4 * 16, ISM, 1, 256, (k_dmma_smem<1, 4><<<grid, 256, sh14>>>(out, ISM, cyc)));
    {
      size_t sh = (16 * 8 * 32 * 4 + 16 * 2 * 8 * 4 * 4) * 8;
      CK(cudaFuncSetAttribute(k_dmma_smem<8, 2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
      RUN("dmma_smem_rt8_ct2_occ1x512_16warps", 16 * 8 * 2 * 16, ISM, 1, 512, (k_dmma_smem<8, 2, 512><<<grid, 512, sh>>>(out, ISM, cyc)));
      size_t sh2 = (16 * 2 * 32 * 4 + 16 * 2 * 8 * 4 * 4) * 8;
      CK(cudaFuncSetAttribute(k_dmma_smem<2, 2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh2));
      RUN("dmma_smem_rt2_ct2_occ1x512_16warps", 16 * 2 * 2 * 16, ISM, 1, 512, (k_dmma_smem<2, 2, 512><<<grid, 512, sh2>>>(out, ISM, cyc)));
      size_t sh1 = (16 * 1 * 32 * 4 + 16 * 2 * 8 * 4 * 4) * 8;
      CK(cudaFuncSetAttribute(k_dmma_smem<1, 2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh1));
      RUN("dmma_smem_rt1_ct2_occ1x512_16warps", 16 * 1 * 2 * 16, ISM, 1, 512, (k_dmma_smem<1, 2, 512><<<grid, 512, sh1>>>(out, ISM, cyc)));
    }
    size_t sh82 = (16 * 8 * 32 * 4 + 16 * 2 * 16 * 4) * 8;
    CK(cudaFuncSetAttribute(k_dmma_smem<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh82));
    RUN("dmma_smem_rt8_ct2_occ1x256", 16 * 8 * 2 * 16, ISM, 1, 256, (k_dmma_smem<8, 2><<<grid, 256, sh82>>>(out, ISM, cyc)));
  }
  return 0;
}

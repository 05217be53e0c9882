// Measured FP64 peak of the device: a register-resident DMMA.8x8x4 loop and a DFMA loop, timed with CUDA events.
// bench.py divides the scoring kernel's algorithmic FLOP/s by the DMMA figure (MEASURED_PEAKS.json carries no
// FP64 entry).  On B200 both pipes share the FP64 units: 64 FMA/clk/SM -> 37.2 TFLOP/s at 1965 MHz.
#include <cuda_runtime.h>

namespace bode {

__global__ void __launch_bounds__(256) k_peak_dmma(double* out, int iters, double x, double y) {
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  const double a = x + threadIdx.x * 1e-9, b = y;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_dfma(double* out, int iters, double x, double y) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t measure_fp64_peak(double* dmma, double* dfma) {
  int dev = 0, sms = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  double* out = nullptr;
  if ((e = cudaMalloc(&out, sizeof(double) * sms * 4 * 256)) != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int iters = 20000;
  float best_m = 1e30f, best_f = 1e30f;
  for (int r = 0; r < 7; r++) {
    cudaEventRecord(a);
    k_peak_dmma<<<sms, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (r >= 2 && ms < best_m) best_m = ms;
    cudaEventRecord(a);
    k_peak_dfma<<<sms * 4, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    if (r >= 2 && ms < best_f) best_f = ms;
  }
  e = cudaGetLastError();
  *dmma = (double)sms * 256 * 8 * 16.0 * iters / (best_m * 1e-3) * 1e-12;
  *dfma = (double)sms * 4 * 256 * 8 * 2.0 * iters / (best_f * 1e-3) * 1e-12;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  return e;
}

}  // namespace bode

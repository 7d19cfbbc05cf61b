// FP64 peak of the whole chip, measured with CUDA events (not clock64): DMMA (mma.sync.m8n8k4.f64, the instruction the
// band factorisation's tile updates use) and plain DFMA, 1024 threads per CTA, 2 CTAs per SM's worth of grid.
// Prints one JSON object; tools/fp64_peak.py stores it as profiles/fp64_peak.json (read by bench.py for the roofline
// of the reduced-system factorisation, because MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <ctime>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters) {
  double d[16]; for (int i = 0; i < 16; ++i) d[i] = 0.0;
  const double a = threadIdx.x * 1e-3, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(d[2 * i], d[2 * i + 1], a, b);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += d[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * m + c;
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int dev = 0, sms = 0, khz = 0; CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int grid = 2 * sms, iters = 1 << 15;
  double* out; CK(cudaMalloc(&out, (size_t)grid * 1024 * sizeof(double)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best_mma = 0, best_fma = 0;
  for (int rep = 0; rep < 6; ++rep) {
    float ms = 0;
    CK(cudaEventRecord(e0)); k_dmma<<<grid, 1024>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double f_mma = (double)grid * 32 /*warps*/ * 8 * iters * (8.0 * 8 * 4 * 2) / (ms * 1e-3) / 1e12;
    CK(cudaEventRecord(e0)); k_dfma<<<grid, 1024>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double f_fma = (double)grid * 1024 * 8 * iters * 2.0 / (ms * 1e-3) / 1e12;
    if (rep > 0) { if (f_mma > best_mma) best_mma = f_mma; if (f_fma > best_fma) best_fma = f_fma; }
  }
  char when[64]; time_t t = time(nullptr); strftime(when, sizeof(when), "%Y-%m-%dT%H:%M:%SZ", gmtime(&t));
  const double ghz = khz * 1e-6;
  printf("{\"fp64_tflops\": %.3f, \"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"sm_count\": %d, \"sm_max_mhz\": %.0f, "
         "\"fma_per_clk_per_sm_at_max_clock\": %.2f, \"gpu_name\": \"%s\", \"when\": \"%s\", "
         "\"how\": \"tools/ubench/fp64_peak.cu: %d CTAs x 1024 threads, %d iterations of 8 independent mma.sync.m8n8k4.f64 (DMMA) / 8 DFMA per thread, "
         "CUDA events, best of 5 after one warm-up; fp64_tflops = max(DMMA, DFMA)\"}\n",
         best_mma > best_fma ? best_mma : best_fma, best_mma, best_fma, sms, khz * 1e-3, (best_mma > best_fma ? best_mma : best_fma) * 1e12 / 2 / sms / (ghz * 1e9), prop.name, when, grid, iters);
  return 0;
}

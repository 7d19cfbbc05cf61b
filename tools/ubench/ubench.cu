// Micro-benchmarks that drive the design of the dense stage (run on the B200 box):
//   DFMA / DMMA(m8n8k4) throughput per SM, SHFL throughput, dependent FP64 reciprocal latency,
//   and the cost of straight-line (fully unrolled) code as a function of its size (instruction cache).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void k_dfma(double* out, int iters, long long* cyc) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double m = 1.0000001, c = 1e-9;
  __syncthreads(); long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * m + c;
  }
  __syncthreads(); long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dmma(double* out, int iters, long long* cyc) {
  double d[8]; for (int i = 0; i < 8; ++i) d[i] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1e-3;
  __syncthreads(); long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    dmma884(d[0], d[1], a, b); dmma884(d[2], d[3], a, b); dmma884(d[4], d[5], a, b); dmma884(d[6], d[7], a, b);
  }
  __syncthreads(); long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl(double* out, int iters, long long* cyc) {
  double a[4]; for (int i = 0; i < 4; ++i) a[i] = threadIdx.x + i;
  __syncthreads(); long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] += __shfl_sync(0xffffffffu, a[(i + 1) & 3], (it + i) & 31);
  }
  __syncthreads(); long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rcp_chain(double* out, int iters, long long* cyc) {
  double a = 1.5 + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = 1.0 / a + 0.25;
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_chain(double* out, int iters, long long* cyc) {
  double a = 1.5 + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = a * 0.999 + 0.25;
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_chain(double* out, int iters, long long* cyc) {
  double a = 1.5 + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_chain(double* out, int iters, long long* cyc) {
  __shared__ int idx[32];
  idx[threadIdx.x] = (threadIdx.x + 1) & 31; __syncwarp();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) j = idx[j];
  long long t1 = clock64();
  out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sts_lds_chain(double* out, int iters, long long* cyc) {
  __shared__ double buf[2][32];
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { buf[it & 1][threadIdx.x] = a; __syncwarp(); a = buf[it & 1][(threadIdx.x + 1) & 31] + 1.0; }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// straight-line code of N independent-ish FMAs (4 accumulators), body repeated by an outer loop
template <int N>
__global__ void k_icache(double* out, int iters, long long* cyc) {
  double a[4] = {1.0 + threadIdx.x, 2.0, 3.0, 4.0};
  const double m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i & 3] = a[i & 3] * m + c;
  }
  long long t1 = clock64();
  out[threadIdx.x] = a[0] + a[1] + a[2] + a[3]; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h[256];
  CK(cudaMalloc(&out, 148 * 1024 * sizeof(double))); CK(cudaMalloc(&cyc, 256 * sizeof(long long)));
  const int it = 4096;
  for (int threads : {128, 256, 512, 1024}) {
    k_dfma<<<148, threads>>>(out, it, cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("DFMA  %4d thr/SM: %.2f FMA/clk/SM\n", threads, (double)threads * 8 * it / h[0]);
    k_dmma<<<148, threads>>>(out, it, cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("DMMA  %4d thr/SM: %.2f FMA/clk/SM  (%.1f cyc per DMMA per SMSP)\n", threads, (double)(threads / 32) * 4 * 256 * it / h[0], (double)h[0] / ((double)(threads / 32) * 4 * it / 4.0));
    k_shfl<<<148, threads>>>(out, it, cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("SHFL  %4d thr/SM: %.3f SHFL.32 warp-instr/clk/SM\n", threads, (double)(threads / 32) * 8 * it / h[0]);
  }
  k_rcp_chain<<<1, 32>>>(out, 1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("dependent (1/x + c) chain: %.1f cyc/iter\n", h[0] / 1000.0);
  k_dfma_chain<<<1, 32>>>(out, 1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("dependent DFMA chain: %.1f cyc/iter\n", h[0] / 1000.0);
  k_shfl_chain<<<1, 32>>>(out, 1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("dependent 64-bit SHFL chain: %.1f cyc/iter\n", h[0] / 1000.0);
  k_lds_chain<<<1, 32>>>(out, 1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("dependent LDS chain: %.1f cyc/iter\n", h[0] / 1000.0);
  k_sts_lds_chain<<<1, 32>>>(out, 1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("STS->syncwarp->LDS->DADD chain: %.1f cyc/iter\n", h[0] / 1000.0);
#define IC(N) { k_icache<N><<<1, 32>>>(out, 64, cyc); CK(cudaDeviceSynchronize()); k_icache<N><<<1, 32>>>(out, 64, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost)); \
    printf("straight-line %5d DFMA (%4d KB code), 1 warp : %.2f cyc/instr\n", N, N * 16 / 1024, h[0] / (64.0 * N)); }
  IC(512) IC(1024) IC(2048) IC(4096) IC(6144) IC(8192) IC(12288)
  return 0;
}

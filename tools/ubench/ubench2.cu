// Isolated timing of the register LDL^T (warp_ldlt32) and the substitution (warp_trsm32) of ba_dense.cuh.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bundleadjustment_benchmarks_b200/csrc/ba_dense.cuh"
using namespace ba;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256, 1) k_fac(double* out, int iters, long long* cyc, int mode) {
  __shared__ double sCol[2][NB];
  __shared__ __align__(16) double sLT[NB][NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < NB * NB; i += blockDim.x) sLT[i >> 5][i & 31] = 0.001 * ((i * 7) % 13);
  __syncthreads();
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (warp == 0) {
      double a[NB]; double z = lane;
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = (c == lane) ? 40.0 + it : 1.0 / (1 + c + lane);
      if (mode == 0) warp_ldlt32<double>(a, z, lane, sCol); else warp_trsm32<double>(a, sLT);
#pragma unroll
      for (int c = 0; c < NB; ++c) acc += a[c];
      acc += z;
    }
    if (blockDim.x > 32) __syncthreads();
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc; long long h[16];
  CK(cudaMalloc(&out, 1024 * sizeof(double))); CK(cudaMalloc(&cyc, 64 * sizeof(long long)));
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {32, 256}) {
      k_fac<<<1, threads>>>(out, 200, cyc, mode); CK(cudaDeviceSynchronize());
      k_fac<<<1, threads>>>(out, 200, cyc, mode); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
      printf("%s, %3d threads in CTA: %.0f cycles per call\n", mode ? "warp_trsm32" : "warp_ldlt32", threads, h[0] / 200.0);
    }
  return 0;
}

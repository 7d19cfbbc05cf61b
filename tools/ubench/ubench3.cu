// L2-resident load latency/throughput by load flavour, and cluster.sync cost.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE> __device__ __forceinline__ unsigned long long ld(const unsigned long long* p) {
  unsigned long long v;
  if (MODE == 0) asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == 1) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == 2) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == 3) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// pointer chase: buf[i] holds the index of the next element (stride of 4 KB apart), L2 resident
template <int MODE> __global__ void k_chase(const unsigned long long* buf, int iters, long long* cyc, unsigned long long* out) {
  unsigned long long j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) j = ld<MODE>(buf + j);
  long long t1 = clock64();
  out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// 32 independent loads per thread (one tile row each lane), then sum
template <int MODE> __global__ void k_batch(const unsigned long long* buf, int iters, long long* cyc, unsigned long long* out) {
  unsigned long long s = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    unsigned long long v[32];
    const unsigned long long* p = buf + (size_t)((i * 32 + threadIdx.x) & 4095) * 549;
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = ld<MODE>(p + c);
#pragma unroll
    for (int c = 0; c < 32; ++c) s += v[c];
  }
  long long t1 = clock64();
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void __cluster_dims__(16, 1, 1) k_csync(int iters, long long* cyc) {
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) cluster.sync();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void __cluster_dims__(16, 1, 1) k_csync_split(int iters, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void __cluster_dims__(16, 1, 1) k_csync_relaxed(int iters, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_syncthreads(int iters, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  const size_t N = 4096 * 549 + 64;  // 18 MB: L2 resident
  unsigned long long *buf, *out; long long* cyc; long long h;
  CK(cudaMalloc(&buf, N * 8)); CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMalloc(&cyc, 64));
  unsigned long long* hb = (unsigned long long*)malloc(N * 8);
  for (size_t i = 0; i < N; ++i) hb[i] = (i + 549 * 7 + 3) % (4096 * 549);
  CK(cudaMemcpy(buf, hb, N * 8, cudaMemcpyHostToDevice));
  const char* names[] = {"ld.global.ca", "ld.global.cg", "ld.volatile.global", "ld.relaxed.gpu", "ld.global.cv"};
#define RUN(M) { k_chase<M><<<1, 32>>>(buf, 2000, cyc, out); CK(cudaDeviceSynchronize()); k_chase<M><<<1, 32>>>(buf, 2000, cyc, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
    printf("%-20s dependent chain: %.0f cyc/load", names[M], h / 2000.0); \
    k_batch<M><<<1, 32>>>(buf, 100, cyc, out); CK(cudaDeviceSynchronize()); k_batch<M><<<1, 32>>>(buf, 100, cyc, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
    printf("   32 independent loads/lane (row-strided): %.0f cyc per batch\n", h / 100.0); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
  for (int threads : {256, 512}) {
    k_csync<<<16, threads>>>(1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("cluster.sync (16 CTAs x %d thr): %.0f cyc\n", threads, h / 1000.0);
    k_csync_split<<<16, threads>>>(1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("barrier.cluster arrive.release + wait.acquire: %.0f cyc\n", h / 1000.0);
    k_csync_relaxed<<<16, threads>>>(1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("barrier.cluster arrive.relaxed + wait: %.0f cyc\n", h / 1000.0);
    k_syncthreads<<<1, threads>>>(1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("__syncthreads (%d thr): %.0f cyc\n", threads, h / 1000.0);
  }
  return 0;
}

// Reduced camera block on sm_100a: blocked LDL^T of the symmetric (block-)banded matrix S and the
// triangular solves. Replaces the right-block solver of the reference,
//   SimplicialLDLT<JacobianType, Lower>::compute / solve      BacktrackLevMarqQRChol.h:339-341,
//                                                              BacktrackLevMarqCholesky.h:278-282
// (stock Eigen, NOT IN TREE). Un-pivoted LDL^T like SimplicialLDLT (D may be negative), natural
// camera order; the band (co-visibility window) plays the role of the sparsity pattern.
//
// Storage: lower band, entry (i,j), 0 <= i-j <= kd, at v[i*lds + j] (lds = kd, v = base + kd), i.e.
// LAPACK 'L' band storage viewed row-wise, so a 32x32 tile is an ordinary strided matrix.
// One cooperative persistent kernel: per 32-column panel every CTA refactors the 32x32 diagonal tile
// redundantly in shared memory (no broadcast needed), CTAs split the triangular solves of the row
// tiles, grid.sync, CTAs split the trailing tile updates, grid.sync.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace ba {
namespace cg = cooperative_groups;

constexpr int NB = 32;
constexpr int DENSE_THREADS = 256;

template <class T> struct BandMat { T* v; size_t lds; int n; int kd; };

template <class T> __device__ __forceinline__ bool band_ok(const BandMat<T>& A, int i, int j) {
  return i < A.n && j <= i && (i - j) <= A.kd;
}

template <class T>
__global__ void __launch_bounds__(DENSE_THREADS) k_band_ldlt(BandMat<T> A, T* __restrict__ dvec, int* __restrict__ info) {
  cg::grid_group grid = cg::this_grid();
  __shared__ T sD[NB][NB + 1];
  __shared__ T sW[NB][NB + 1];
  __shared__ T sA[NB][NB + 1];
  __shared__ T sB[NB][NB + 1];
  __shared__ T sd[NB];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;  // 8 warps
  const int n = A.n, kd = A.kd;
  const int nt = (n + NB - 1) / NB;
  const int bt = (kd + NB - 1) / NB;  // row tiles below the diagonal tile that a panel can touch
  for (int k = 0; k < nt; ++k) {
    const int k0 = k * NB;
    for (int r = ty; r < NB; r += 8) {
      const int gi = k0 + r, gj = k0 + tx;
      sD[r][tx] = (gi < n) ? (band_ok(A, gi, gj) ? A.v[(size_t)gi * A.lds + gj] : T(0)) : (r == tx ? T(1) : T(0));
    }
    // un-blocked right-looking LDL^T of the tile; column scaling deferred
    for (int j = 0; j < NB; ++j) {
      __syncthreads();
      const T dj = sD[j][j];
      const T inv = T(1) / dj;
      for (int i = j + 1 + ty; i < NB; i += 8) {
        if (tx > j && tx <= i) sD[i][tx] -= sD[i][j] * sD[tx][j] * inv;
      }
    }
    __syncthreads();
    if (ty == 0) {
      const T d = sD[tx][tx];
      sd[tx] = d;
      if (blockIdx.x == 0 && k0 + tx < n && (d == T(0) || !(d == d))) atomicCAS(info, 0, k0 + tx + 1);
    }
    __syncthreads();
    for (int r = ty; r < NB; r += 8) if (tx < r) sD[r][tx] = sD[r][tx] / sd[tx];
    __syncthreads();
    // W = inverse of the unit lower L_kk; lane tx builds column tx by forward substitution
    if (ty == 0) {
      T x[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        T s = (i == tx) ? T(1) : T(0);
#pragma unroll
        for (int m = 0; m < i; ++m) s -= (m >= tx) ? sD[i][m] * x[m] : T(0);
        x[i] = (i >= tx) ? s : T(0);
        sW[i][tx] = x[i];
      }
    }
    __syncthreads();
    // triangular solves of the row tiles: L_ik = A_ik * L_kk^-T * D^-1
    const int last = min(nt - 1, k + bt);
    for (int it = k + 1 + blockIdx.x; it <= last; it += gridDim.x) {
      for (int r = ty; r < NB; r += 8) {
        const int gi = it * NB + r, gj = k0 + tx;
        sA[r][tx] = band_ok(A, gi, gj) ? A.v[(size_t)gi * A.lds + gj] : T(0);
      }
      __syncthreads();
      T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll 8
      for (int m = 0; m < NB; ++m) {
        const T w = sW[tx][m];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += sA[ty * 4 + q][m] * w;
      }
      const T idc = T(1) / sd[tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = it * NB + ty * 4 + q, gj = k0 + tx;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] = acc[q] * idc;
      }
      __syncthreads();
    }
    grid.sync();
    if (blockIdx.x == 0) {  // write back the factored diagonal tile (nobody reads it again here)
      for (int r = ty; r < NB; r += 8) {
        const int gi = k0 + r, gj = k0 + tx;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] = (r == tx) ? sd[tx] : sD[r][tx];
      }
      if (ty == 0 && k0 + tx < n) dvec[k0 + tx] = sd[tx];
    }
    // trailing update A_ij -= L_ik D_k L_jk^T for k < j <= i <= last
    const int nb = last - k;
    const int npairs = nb * (nb + 1) / 2;
    for (int pidx = blockIdx.x; pidx < npairs; pidx += gridDim.x) {
      int ii = (int)((sqrtf(8.0f * (float)pidx + 1.0f) - 1.0f) * 0.5f);
      while ((ii + 1) * (ii + 2) / 2 <= pidx) ++ii;
      while (ii * (ii + 1) / 2 > pidx) --ii;
      const int jj = pidx - ii * (ii + 1) / 2;
      const int it = k + 1 + ii, jt = k + 1 + jj;
      if ((it - jt) * NB - (NB - 1) > kd) continue;  // tile entirely outside the band
      for (int r = ty; r < NB; r += 8) {
        const int gi = it * NB + r, gj2 = jt * NB + r, gc = k0 + tx;
        sA[r][tx] = band_ok(A, gi, gc) ? A.v[(size_t)gi * A.lds + gc] : T(0);
        sB[r][tx] = band_ok(A, gj2, gc) ? A.v[(size_t)gj2 * A.lds + gc] * sd[tx] : T(0);
      }
      __syncthreads();
      T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll 8
      for (int m = 0; m < NB; ++m) {
        const T b = sB[tx][m];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += sA[ty * 4 + q][m] * b;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = it * NB + ty * 4 + q, gj = jt * NB + tx;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] -= acc[q];
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// y = S^-1 g from the factor above: L z = g, w = D^-1 z, L^T y = w. Single CTA, 16 warps; the
// 32x32 diagonal systems are solved by warp 0 with shuffles, the band updates are split over warps.
constexpr int SOLVE_THREADS = 512;

template <class T>
__global__ void __launch_bounds__(SOLVE_THREADS) k_band_ldlt_solve(BandMat<T> A, const T* __restrict__ dvec, const T* __restrict__ g,
                                                                   T* __restrict__ y, T sign) {
  __shared__ T sL[NB][NB + 1];
  __shared__ T sb[NB];
  __shared__ T sacc[SOLVE_THREADS / 32][NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = SOLVE_THREADS / 32;
  const int n = A.n, kd = A.kd, nt = (n + NB - 1) / NB;
  for (int i = tid; i < n; i += SOLVE_THREADS) y[i] = g[i];
  __syncthreads();
  // forward
  for (int k = 0; k < nt; ++k) {
    const int k0 = k * NB;
    for (int r = warp; r < NB; r += nw) {
      const int gi = k0 + r, gj = k0 + lane;
      sL[r][lane] = (gj < gi && band_ok(A, gi, gj)) ? A.v[(size_t)gi * A.lds + gj] : T(0);
    }
    __syncthreads();
    if (warp == 0) {
      T b = (k0 + lane < n) ? y[k0 + lane] : T(0);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const T bj = __shfl_sync(0xffffffffu, b, j);
        if (lane > j) b -= sL[lane][j] * bj;
      }
      sb[lane] = b;
      if (k0 + lane < n) y[k0 + lane] = b;
    }
    __syncthreads();
    const int r1 = min(n - 1, k0 + NB - 1 + kd);
    for (int r = k0 + NB + warp; r <= r1; r += nw) {
      const int gj = k0 + lane;
      T v = band_ok(A, r, gj) ? A.v[(size_t)r * A.lds + gj] * sb[lane] : T(0);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) y[r] -= v;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += SOLVE_THREADS) y[i] = y[i] / dvec[i];
  __syncthreads();
  // backward
  for (int k = nt - 1; k >= 0; --k) {
    const int k0 = k * NB;
    for (int r = warp; r < NB; r += nw) {
      const int gi = k0 + r, gj = k0 + lane;
      sL[r][lane] = (gj < gi && band_ok(A, gi, gj)) ? A.v[(size_t)gi * A.lds + gj] : T(0);
    }
    T acc = T(0);
    const int r1 = min(n - 1, k0 + NB - 1 + kd);
    for (int r = k0 + NB + warp; r <= r1; r += nw) {
      const int gj = k0 + lane;
      if (band_ok(A, r, gj)) acc += A.v[(size_t)r * A.lds + gj] * y[r];
    }
    sacc[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      T b = (k0 + lane < n) ? y[k0 + lane] : T(0);
      for (int w = 0; w < nw; ++w) b -= sacc[w][lane];
#pragma unroll
      for (int j = NB - 1; j >= 0; --j) {
        const T bj = __shfl_sync(0xffffffffu, b, j);
        if (lane < j) b -= sL[j][lane] * bj;
      }
      if (k0 + lane < n) y[k0 + lane] = b;
    }
    __syncthreads();
  }
  if (sign != T(1)) { for (int i = tid; i < n; i += SOLVE_THREADS) y[i] = sign * y[i]; }
}

// ---------------------------------------------------------------------------------------------
// Cluster-resident variant: ONE thread-block cluster (8 portable / 16 CTAs) factors the band matrix.
// The per-panel work of a banded factorisation is tiny (<= bt(bt+1)/2 32^3 tile updates), so the
// chain of n/32 dependent panels is latency-bound: cluster barriers (~0.2 us) replace grid-wide
// barriers (~4 us), the forward substitution L z = g rides along inside the factorisation (the
// right-hand side is one more row of the matrix), W_k = L_kk^-1 is produced by the same 32
// elimination steps as the tile factor and kept for the backward pass, which CTA 0 runs alone.
// ---------------------------------------------------------------------------------------------
constexpr int CL_THREADS = 512;
constexpr int CL_GROUPS = CL_THREADS / 128;

template <class T> struct ClusterSmem {
  T sD[NB][NB + 1]; T sW[NB][NB + 1]; T sd[NB]; T sz[NB];
  T gA[CL_GROUPS][NB][NB + 1]; T gB[CL_GROUPS][NB][NB + 1];
};

__device__ __forceinline__ void group_barrier(int group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(128) : "memory"); }

// batched, unconditional tile loads: 8 independent L2 requests per lane in flight
template <class T>
__device__ __forceinline__ void tile_fetch(const BandMat<T>& A, int row0, int col0, int gt, T (&reg)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int idx = gt + 128 * q, r = idx >> 5, c = idx & 31, gi = row0 + r, gj = col0 + c;
    const bool ok = band_ok(A, gi, gj);
    const T* p = ok ? (A.v + (size_t)gi * A.lds + gj) : A.v;
    const T v = *p;
    reg[q] = ok ? v : T(0);
  }
}
template <class T>
__device__ __forceinline__ void tile_stage(T (&dst)[NB][NB + 1], int gt, const T (&reg)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) { const int idx = gt + 128 * q; dst[idx >> 5][idx & 31] = reg[q]; }
}

template <class T>
__global__ void __launch_bounds__(CL_THREADS, 1) k_band_ldlt_cluster(BandMat<T> A, T* __restrict__ dvec, T* __restrict__ Wbuf,
                                                                      T* __restrict__ rhs, T* __restrict__ y, T sign, int* __restrict__ info, long long* __restrict__ dbg) {
  cg::cluster_group cluster = cg::this_cluster();
  long long tc[6] = {0, 0, 0, 0, 0, 0}; long long t0 = clock64();
#define TICK(i) { const long long t1_ = clock64(); tc[i] += t1_ - t0; t0 = t1_; }
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) unsigned char cl_smem_raw[];
  ClusterSmem<T>& sm = *reinterpret_cast<ClusterSmem<T>*>(cl_smem_raw);
  T(&sD)[NB][NB + 1] = sm.sD;
  T(&sW)[NB][NB + 1] = sm.sW;
  T(&sd)[NB] = sm.sd;
  T(&sz)[NB] = sm.sz;
  T(&gA)[CL_GROUPS][NB][NB + 1] = sm.gA;
  T(&gB)[CL_GROUPS][NB][NB + 1] = sm.gB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, group = warp >> 2, gt = tid & 127;
  const int n = A.n, kd = A.kd;
  const int nt = (n + NB - 1) / NB, bt = (kd + NB - 1) / NB;
  const int gstride = CL_GROUPS * C, gfirst = rank * CL_GROUPS + group;
  for (int k = 0; k < nt; ++k) {
    const int k0 = k * NB;
    const int last = min(nt - 1, k + bt);
    // prefetch this group's first row tile of the panel while the diagonal tile is being factored
    T ra[8], rb[8], rc[8];
    int it = k + 1 + gfirst;
    if (it <= last) tile_fetch<T>(A, it * NB, k0, gt, ra);
    for (int idx = tid; idx < NB * NB; idx += CL_THREADS) {
      const int r = idx >> 5, c = idx & 31, gi = k0 + r, gj = k0 + c;
      sD[r][c] = (gi < n) ? (band_ok(A, gi, gj) ? A.v[(size_t)gi * A.lds + gj] : T(0)) : (r == c ? T(1) : T(0));
      sW[r][c] = (r == c) ? T(1) : T(0);
    }
    if (tid < NB) sz[tid] = (k0 + tid < n) ? rhs[k0 + tid] : T(0);
    // 32 elimination steps on [A_kk | I | g_k]: A_kk -> L D (column scaling deferred), I -> L^-1, g_k -> z_k
    for (int j = 0; j < NB; ++j) {
      __syncthreads();
      const T inv = T(1) / sD[j][j];
      for (int idx = tid; idx < NB * NB; idx += CL_THREADS) {
        const int i = idx >> 5, c = idx & 31;
        if (i > j) {
          const T l = sD[i][j] * inv;
          if (c > j) { if (c <= i) sD[i][c] -= l * sD[c][j]; }
          else sW[i][c] -= l * sW[j][c];
        }
      }
      if (tid < NB && tid > j) sz[tid] -= sD[tid][j] * inv * sz[j];
    }
    __syncthreads();
    if (tid < NB) {
      const T d = sD[tid][tid];
      sd[tid] = d;
      if (rank == 0 && k0 + tid < n && (d == T(0) || !(d == d))) atomicCAS(info, 0, k0 + tid + 1);
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += CL_THREADS) { const int r = idx >> 5, c = idx & 31; if (c < r) sD[r][c] = sD[r][c] / sd[c]; }
    TICK(0)
    // row tiles below: L_ik = A_ik W^T D^-1 ; forward substitution g_i -= L_ik z_k
    for (; it <= last; it += gstride) {
      tile_stage<T>(gA[group], gt, ra);
      group_barrier(group);
      if (it + gstride <= last) tile_fetch<T>(A, (it + gstride) * NB, k0, gt, ra);
      const int c = gt & 31, rb8 = (gt >> 5) * 8;
      T acc[8] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
#pragma unroll 4
      for (int m = 0; m < NB; ++m) {
        const T w = sW[c][m];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += gA[group][rb8 + q][m] * w;
      }
      const T idc = T(1) / sd[c];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int gi = it * NB + rb8 + q, gj = k0 + c;
        const bool ok = band_ok(A, gi, gj);
        const T v = ok ? acc[q] * idc : T(0);
        gB[group][rb8 + q][c] = v;
        if (ok) A.v[(size_t)gi * A.lds + gj] = v;
      }
      group_barrier(group);
      if (gt < NB) {
        const int gi = it * NB + gt;
        if (gi < n) {
          T s = T(0);
#pragma unroll 8
          for (int cc = 0; cc < NB; ++cc) s += gB[group][gt][cc] * sz[cc];
          rhs[gi] -= s;
        }
      }
      group_barrier(group);
    }
    TICK(1)
    cluster.sync();
    TICK(2)
    // trailing update A_ij -= L_ik D_k L_jk^T, k < j <= i <= last; operands one tile ahead in registers
    const int nb = last - k;
    const int npairs = nb * (nb + 1) / 2;
    auto decode = [&](int pidx, int& ti, int& tj) {
      int ii = (int)((sqrtf(8.0f * (float)pidx + 1.0f) - 1.0f) * 0.5f);
      while ((ii + 1) * (ii + 2) / 2 <= pidx) ++ii;
      while (ii * (ii + 1) / 2 > pidx) --ii;
      ti = k + 1 + ii; tj = k + 1 + (pidx - ii * (ii + 1) / 2);
    };
    int pidx = gfirst, ti = 0, tj = 0;
    if (pidx < npairs) { decode(pidx, ti, tj); tile_fetch<T>(A, ti * NB, k0, gt, ra); tile_fetch<T>(A, tj * NB, k0, gt, rb); tile_fetch<T>(A, ti * NB, tj * NB, gt, rc); }
    if (rank == 0) {  // publish the factored diagonal tile, D, W_k and w_k = D^-1 z_k (overlaps the fetch latency)
      for (int idx = tid; idx < NB * NB; idx += CL_THREADS) {
        const int r = idx >> 5, c = idx & 31, gi = k0 + r, gj = k0 + c;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] = (r == c) ? sd[c] : sD[r][c];
        Wbuf[(size_t)k * NB * NB + idx] = sW[r][c];
      }
      if (tid < NB && k0 + tid < n) { dvec[k0 + tid] = sd[tid]; rhs[k0 + tid] = sz[tid] / sd[tid]; }
    }
    for (; pidx < npairs; pidx += gstride) {
      const int ci = ti, cj = tj;
      tile_stage<T>(gA[group], gt, ra);
#pragma unroll
      for (int q = 0; q < 8; ++q) { const int idx = gt + 128 * q; gB[group][idx >> 5][idx & 31] = rb[q] * sd[idx & 31]; }
      T cc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) cc[q] = rc[q];
      group_barrier(group);
      if (pidx + gstride < npairs) { decode(pidx + gstride, ti, tj); tile_fetch<T>(A, ti * NB, k0, gt, ra); tile_fetch<T>(A, tj * NB, k0, gt, rb); tile_fetch<T>(A, ti * NB, tj * NB, gt, rc); }
      // lane layout of the fetch: element idx = gt + 128 q -> row (idx>>5), col (idx&31): compute the same elements
      T acc[8] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
      const int c = gt & 31, r0 = gt >> 5;  // rows r0, r0+4, ..., r0+28
#pragma unroll 4
      for (int m = 0; m < NB; ++m) {
        const T b = gB[group][c][m];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += gA[group][r0 + 4 * q][m] * b;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int gi = ci * NB + r0 + 4 * q, gj = cj * NB + c;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] = cc[q] - acc[q];
      }
      group_barrier(group);
    }
    TICK(3)
    cluster.sync();
    TICK(4)
  }
  if (rank != 0) return;
  // ------------------------------------------------------------------ backward pass on CTA 0
  // y_k = W_k^T (w_k - sum_{r > k0+31} L[r][k0+c] y[r]). Software pipeline: warps 1..15 accumulate, one
  // step ahead, the rows that are already final (r >= k0+64); warp 0 adds the tile (k+1,k) product with
  // the just-computed y_{k+1} and applies W_k^T. One block barrier per step.
  T(*sacc)[NB] = reinterpret_cast<T(*)[NB]>(&gA[0][0][0]);          // [2][16][32] partial sums, double-buffered
  T(*sLk)[NB + 1] = reinterpret_cast<T(*)[NB + 1]>(&gB[0][0][0]);   // [2][32][33] tile (k+1,k), double-buffered
  T(*sWk)[NB + 1] = reinterpret_cast<T(*)[NB + 1]>(&gB[2][0][0]);   // [2][32][33] W_k, double-buffered
  __shared__ T sy[2][NB];
  constexpr int NWARP = CL_THREADS / 32;
  auto stage = [&](int k, int buf) {  // executed by warps 1..15 for step k
    const int k0 = k * NB, w = warp - 1;
    T acc = T(0);
    const int r1 = min(n - 1, k0 + NB - 1 + kd), gj = k0 + lane;
    if (gj < n) {
#pragma unroll 4
      for (int r = k0 + 2 * NB + w; r <= r1; r += NWARP - 1)
        if (r - gj <= kd) acc += A.v[(size_t)r * A.lds + gj] * y[r];
    }
    sacc[buf * NWARP + warp][lane] = acc;
    for (int idx = (warp - 1) * 32 + lane; idx < NB * NB; idx += (NWARP - 1) * 32) {
      const int r = idx >> 5, c = idx & 31, gi = k0 + NB + r, gjj = k0 + c;
      sLk[buf * NB + r][c] = (gi < n && gjj < n && gi - gjj <= kd) ? A.v[(size_t)gi * A.lds + gjj] : T(0);
      sWk[buf * NB + r][c] = Wbuf[(size_t)k * NB * NB + idx];
    }
  };
  if (tid < NB) { sy[0][tid] = T(0); sy[1][tid] = T(0); }
  if (warp > 0) stage(nt - 1, (nt - 1) & 1);
  __syncthreads();
  for (int k = nt - 1; k >= 0; --k) {
    const int k0 = k * NB, buf = k & 1;
    if (warp == 0) {
      T b = (k0 + lane < n) ? rhs[k0 + lane] : T(0);
#pragma unroll
      for (int w = 1; w < NWARP; ++w) b -= sacc[buf * NWARP + w][lane];
      // tile (k+1,k): rows of tile k+1 times y_{k+1} (kept in sy[(k+1)&1])
      T t = T(0);
#pragma unroll 8
      for (int r = 0; r < NB; ++r) t += sLk[buf * NB + r][lane] * sy[(k + 1) & 1][r];
      if (k + 1 < nt) b -= t;
      T yv = T(0);
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        const T bm = __shfl_sync(0xffffffffu, b, m);
        if (m >= lane) yv += sWk[buf * NB + m][lane] * bm;
      }
      sy[buf][lane] = (k0 + lane < n) ? yv : T(0);
      if (k0 + lane < n) y[k0 + lane] = yv;
    } else if (k > 0) {
      stage(k - 1, (k - 1) & 1);  // needs y rows >= (k-1)*32+64 = k0+32: final since the previous step
    }
    __syncthreads();
  }
  if (sign != T(1)) for (int i = tid; i < n; i += CL_THREADS) y[i] = sign * y[i];
  TICK(5)
  if (dbg && tid == 0) for (int i = 0; i < 6; ++i) dbg[i] = tc[i];
#undef TICK
}

}  // namespace ba

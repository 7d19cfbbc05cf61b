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
      sD[r][tx] = (gi < n && gj <= gi) ? A.v[(size_t)gi * A.lds + gj] : (r == tx ? T(1) : T(0));
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
        if (gi < n && gj <= gi) A.v[(size_t)gi * A.lds + gj] = (r == tx) ? sd[tx] : sD[r][tx];
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
      sL[r][lane] = (gi < n && gj < gi) ? A.v[(size_t)gi * A.lds + gj] : T(0);
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
      sL[r][lane] = (gi < n && gj < gi) ? A.v[(size_t)gi * A.lds + gj] : T(0);
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

}  // namespace ba

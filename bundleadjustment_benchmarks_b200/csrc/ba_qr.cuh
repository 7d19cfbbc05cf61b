// Householder QR of the reduced camera block on sm_100a (QRKIT / MOREQR right block).
// Replaces DenseBlockedThinQR<...,4,true>::compute + matrixQ().transpose()*rhs + the upper
// triangular solve (src/Optimization/BAFunctor.h:101,111; BacktrackLevMarqMore.h:328-344;
// solver NOT IN TREE). Deliberate restatement: the reference factors the tall
// J2bot ((2K+9N) x 9N, dense: 1.3 TB at the synthetic scale); here the point stage accumulates the
// square reduced camera matrix S = J2bot^T J2bot and this kernel factors S = Q R, y = R^-1 Q^T g.
//
// Storage: general band, COLUMN-major so that Householder columns are contiguous:
//   entry (i,j), j-ku <= i <= j+kd, at G[j*ld + (i - j + ku)], ld = kd + ku + 1, ku = min(n-1, 2kd).
// One cooperative persistent kernel; per panel of QR_PB columns CTA 0 factors the panel, grid.sync,
// then one WARP per trailing column applies the panel's reflectors (lanes own fixed rows -> no
// cross-lane hazards), the right-hand side rides along as one more column, grid.sync.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "ba_dense.cuh"

namespace ba {

constexpr int QR_THREADS = 256;
constexpr int QR_PB = 8;
constexpr int QR_SOLVE_THREADS = 512;

template <class T> struct QRMat { T* G; size_t ld; int n; int kd; int ku; };

template <class T> __device__ __forceinline__ T& gq(const QRMat<T>& Q, int i, int j) { return Q.G[(size_t)j * Q.ld + (i - j + Q.ku)]; }

// symmetric lower band -> general band (both triangles), rhs <- g
template <class T>
__global__ void k_band_expand(BandMat<T> A, T* __restrict__ G, size_t ld, int ku, const T* __restrict__ g, T* __restrict__ rhs) {
  const size_t total = (size_t)A.n * ld, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int j = (int)(idx / ld), off = (int)(idx - (size_t)j * ld);
    const int i = j + off - ku;
    T v = T(0);
    if (i >= 0 && i < A.n) {
      const int hi = max(i, j), lo = min(i, j);
      if (hi - lo <= A.kd) v = A.v[(size_t)hi * A.lds + lo];
    }
    G[idx] = v;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)A.n; i += stride) rhs[i] = g[i];
}

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// apply reflector `col` (v in column col of G below the diagonal, tau) to the vector x(r), r in
// [col, rend], given through a pointer xp such that x(r) = xp[r]; lanes own rows r = base+lane+32t.
template <class T>
__device__ __forceinline__ void apply_reflector(const QRMat<T>& Q, int col, int rend, T tau, T* __restrict__ xp, int base, int lane) {
  if (tau == T(0)) return;
  const T* vp = Q.G + (size_t)col * Q.ld + (Q.ku - col);  // v(r) = vp[r]
  T s = T(0);
  int r0 = base + lane;
  while (r0 <= col) r0 += 32;
  for (int r = r0; r <= rend; r += 32) s += vp[r] * xp[r];
  s = warp_sum(s);
  s = (s + xp[col]) * tau;
  __syncwarp();
  for (int r = r0; r <= rend; r += 32) xp[r] -= s * vp[r];
  if (((col - base) & 31) == lane) xp[col] -= s;
  __syncwarp();
}

template <class T>
__global__ void __launch_bounds__(QR_THREADS) k_band_qr(QRMat<T> Q, T* __restrict__ tauv, T* __restrict__ rhs) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ T sred[QR_THREADS / 32];
  __shared__ T sbc[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = QR_THREADS / 32;
  const int n = Q.n, kd = Q.kd, ku = Q.ku;
  const int gwarp = blockIdx.x * nw + warp, nwarps = gridDim.x * nw;
  for (int k0 = 0; k0 < n; k0 += QR_PB) {
    const int pb = min(QR_PB, n - k0);
    if (blockIdx.x == 0) {
      for (int c = 0; c < pb; ++c) {
        const int col = k0 + c, rend = min(n - 1, col + kd);
        T* xp = Q.G + (size_t)col * Q.ld + (ku - col);  // x(r) = xp[r]
        T s = T(0);
        for (int r = col + 1 + tid; r <= rend; r += QR_THREADS) { const T v = xp[r]; s += v * v; }
        s = warp_sum(s);
        if (lane == 0) sred[warp] = s;
        __syncthreads();
        if (tid == 0) {
          T tail2 = T(0);
          for (int w = 0; w < nw; ++w) tail2 += sred[w];
          const T c0 = xp[col];
          T tau = T(0), inv = T(0);
          if (tail2 > (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f))) {
            T beta = sqrt(c0 * c0 + tail2);
            if (c0 >= T(0)) beta = -beta;
            inv = T(1) / (c0 - beta);
            tau = (beta - c0) / beta;
            xp[col] = beta;
          }
          tauv[col] = tau;
          sbc[0] = tau; sbc[1] = inv;
        }
        __syncthreads();
        const T tau = sbc[0], inv = sbc[1];
        if (tau != T(0)) for (int r = col + 1 + tid; r <= rend; r += QR_THREADS) xp[r] *= inv;
        __syncthreads();
        // remaining panel columns: one warp each
        for (int cc = c + 1 + warp; cc < pb; cc += nw) {
          const int j = k0 + cc;
          T* xj = Q.G + (size_t)j * Q.ld + (ku - j);
          apply_reflector<T>(Q, col, rend, tau, xj, k0, lane);
        }
        __syncthreads();
      }
    }
    grid.sync();
    // trailing columns + rhs
    const int jlast = min(n - 1, k0 + pb - 1 + ku);
    const int ntrail = jlast - (k0 + pb) + 1;  // may be <= 0
    for (int w = gwarp; w < ntrail + 1; w += nwarps) {
      T* xj;
      if (w == ntrail) xj = rhs;
      else { const int j = k0 + pb + w; xj = Q.G + (size_t)j * Q.ld + (ku - j); }
      const int j = (w == ntrail) ? n : (k0 + pb + w);
      for (int c = 0; c < pb; ++c) {
        const int col = k0 + c, rend = min(n - 1, col + kd);
        if (w != ntrail && j > col + ku) continue;
        apply_reflector<T>(Q, col, rend, tauv[col], xj, k0, lane);
      }
    }
    grid.sync();
  }
}

// ---------------------------------------------------------------------------------------------
// Register/shared-memory variant for kd + QR_PB <= 32 * QR_MAXR rows per reflector (the banded case):
// the panel is factored by CTA 0 entirely in shared memory; for the trailing update every CTA stages the
// panel's reflector vectors in shared memory once and each warp keeps its column in REGISTERS while the
// QR_PB reflectors are applied (one load and one store of the column per panel instead of one per
// reflector, reflector vectors from shared memory instead of L2).
// ---------------------------------------------------------------------------------------------
constexpr int QR_MAXR = 20;  // rows per lane held in registers

template <class T>
__global__ void __launch_bounds__(QR_THREADS) k_band_qr_reg(QRMat<T> Q, T* __restrict__ tauv, T* __restrict__ rhs) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char qr_smem_raw[];
  const int n = Q.n, kd = Q.kd, ku = Q.ku;
  const int LV = kd + QR_PB;                       // rows k0 .. k0+LV-1 cover every reflector of a panel
  T* sV = reinterpret_cast<T*>(qr_smem_raw);       // [QR_PB][LV]: panel columns (phase 1) / reflector vectors (phase 2)
  T* stau = sV + QR_PB * LV;                       // [QR_PB]
  T* sred = stau + QR_PB;                          // [QR_THREADS / 32 + 2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = QR_THREADS / 32;
  const int gwarp = blockIdx.x * nw + warp, nwarps = gridDim.x * nw;
  for (int k0 = 0; k0 < n; k0 += QR_PB) {
    const int pb = min(QR_PB, n - k0);
    if (blockIdx.x == 0) {
      // panel columns into shared memory: sV[c][r] = G(k0 + r, k0 + c), rows k0 .. min(n-1, k0+c+kd)
      for (int idx = tid; idx < pb * LV; idx += QR_THREADS) {
        const int c = idx / LV, r = idx - c * LV, i = k0 + r, j = k0 + c;
        sV[idx] = (i < n && i <= j + kd) ? gq(Q, i, j) : T(0);
      }
      __syncthreads();
      for (int c = 0; c < pb; ++c) {
        T* v = sV + c * LV;
        const int rend = min(n - 1, k0 + c + kd) - k0;  // last local row of this column
        T s = T(0);
        for (int r = c + 1 + tid; r <= rend; r += QR_THREADS) s += v[r] * v[r];
        s = warp_sum(s);
        if (lane == 0) sred[warp] = s;
        __syncthreads();
        if (tid == 0) {
          T tail2 = T(0);
          for (int w = 0; w < nw; ++w) tail2 += sred[w];
          const T c0 = v[c];
          T tau = T(0), inv = T(0);
          if (tail2 > (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f))) {
            T beta = sqrt(c0 * c0 + tail2);
            if (c0 >= T(0)) beta = -beta;
            inv = T(1) / (c0 - beta);
            tau = (beta - c0) / beta;
            v[c] = beta;
          }
          stau[c] = tau; sred[nw] = inv;
        }
        __syncthreads();
        const T tau = stau[c], inv = sred[nw];
        if (tau != T(0)) for (int r = c + 1 + tid; r <= rend; r += QR_THREADS) v[r] *= inv;
        __syncthreads();
        for (int cc = c + 1 + warp; cc < pb; cc += nw) {  // remaining panel columns: one warp each
          if (tau == T(0)) continue;
          T* x = sV + cc * LV;
          T d = T(0);
          for (int r = c + 1 + lane; r <= rend; r += 32) d += v[r] * x[r];
          d = (warp_sum(d) + x[c]) * tau;
          __syncwarp();
          for (int r = c + 1 + lane; r <= rend; r += 32) x[r] -= d * v[r];
          if (lane == 0) x[c] -= d;
          __syncwarp();
        }
        __syncthreads();
      }
      for (int idx = tid; idx < pb * LV; idx += QR_THREADS) {
        const int c = idx / LV, r = idx - c * LV, i = k0 + r, j = k0 + c;
        if (i < n && i <= j + kd) gq(Q, i, j) = sV[idx];
      }
      if (tid < pb) tauv[k0 + tid] = stau[tid];
    }
    grid.sync();
    // reflector vectors of this panel -> shared memory (v(col) = 1 implicit; sV[c][r] valid for r > c)
    for (int idx = tid; idx < pb * LV; idx += QR_THREADS) {
      const int c = idx / LV, r = idx - c * LV, i = k0 + r, j = k0 + c;
      sV[idx] = (r > c && i < n && i <= j + kd) ? gq(Q, i, j) : T(0);
    }
    if (tid < pb) stau[tid] = tauv[k0 + tid];
    __syncthreads();
    const int jlast = min(n - 1, k0 + pb - 1 + ku);
    const int ntrail = jlast - (k0 + pb) + 1;  // may be <= 0
    const int rlast = min(n - 1, k0 + pb - 1 + kd);  // last row any reflector of the panel touches
    for (int w = gwarp; w < ntrail + 1; w += nwarps) {
      const int j = (w == ntrail) ? n : (k0 + pb + w);
      T* xp = (w == ntrail) ? rhs : (Q.G + (size_t)j * Q.ld + (ku - j));  // x(r) = xp[r]
      const int rlo = (w == ntrail) ? k0 : max(k0, j - ku);  // rows above j - ku are not part of column j's band
      T x[QR_MAXR];
#pragma unroll
      for (int t = 0; t < QR_MAXR; ++t) { const int i = k0 + lane + 32 * t; x[t] = (i >= rlo && i <= rlast) ? xp[i] : T(0); }
      for (int c = 0; c < pb; ++c) {
        const int col = k0 + c;
        const T tau = stau[c];
        if ((w != ntrail && j > col + ku) || tau == T(0)) continue;
        const T* v = sV + c * LV;
        T d = T(0);
#pragma unroll
        for (int t = 0; t < QR_MAXR; ++t) { const int r = lane + 32 * t; if (r < LV) d += v[r] * x[t]; }  // v[r] = 0 for r <= c and beyond the column
        d = warp_sum(d);
        // x(col): row local index c (< QR_PB <= 32) lives in lane c, register 0
        const T xcol = __shfl_sync(0xffffffffu, x[0], c);
        d = (d + xcol) * tau;
#pragma unroll
        for (int t = 0; t < QR_MAXR; ++t) { const int r = lane + 32 * t; if (r < LV) x[t] -= d * v[r]; }
        if (lane == c) x[0] -= d;
      }
#pragma unroll
      for (int t = 0; t < QR_MAXR; ++t) { const int i = k0 + lane + 32 * t; if (i >= rlo && i <= rlast) xp[i] = x[t]; }
    }
    grid.sync();
  }
}

// y = sign * R^-1 (Q^T g): blocked upper-triangular back substitution, single CTA.
template <class T>
__global__ void __launch_bounds__(QR_SOLVE_THREADS) k_band_qr_backsolve(QRMat<T> Q, T* __restrict__ rhs, T* __restrict__ y, T sign) {
  __shared__ T sR[NB][NB + 1];
  __shared__ T sy[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = QR_SOLVE_THREADS / 32;
  const int n = Q.n, ku = Q.ku, nt = (n + NB - 1) / NB;
  for (int k = nt - 1; k >= 0; --k) {
    const int k0 = k * NB;
    for (int c = warp; c < NB; c += nw) {  // sR[r][c] = R(k0+r, k0+c), r <= c
      const int gi = k0 + lane, gj = k0 + c;
      sR[lane][c] = (gj < n && gi <= gj && gj - gi <= ku) ? gq(Q, gi, gj) : ((lane == c) ? T(1) : T(0));
    }
    __syncthreads();
    if (warp == 0) {
      T b = (k0 + lane < n) ? rhs[k0 + lane] : T(0);
#pragma unroll
      for (int j = NB - 1; j >= 0; --j) {
        const T yj = __shfl_sync(0xffffffffu, b, j) / sR[j][j];
        if (lane == j) b = yj;
        if (lane < j) b -= sR[lane][j] * yj;
      }
      sy[lane] = b;
      if (k0 + lane < n) y[k0 + lane] = sign * b;
    }
    __syncthreads();
    const int ilo = max(0, k0 - ku);
    for (int i = ilo + tid; i < k0; i += QR_SOLVE_THREADS) {
      T acc = T(0);
#pragma unroll 8
      for (int c = 0; c < NB; ++c) {
        const int gj = k0 + c;
        if (gj < n && gj - i <= ku) acc += gq(Q, i, gj) * sy[c];
      }
      rhs[i] -= acc;
    }
    __syncthreads();
  }
}

}  // namespace ba

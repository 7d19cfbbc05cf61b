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
constexpr int QR_RPT = (32 * QR_MAXR + QR_THREADS - 1) / QR_THREADS;  // panel rows per thread in the panel factorisation

// Look-ahead: CTA 0 is the PANEL CTA. In iteration k it gives the columns of panel k+1 panel k's update and
// factors panel k+1 in shared memory while every other CTA updates the remaining trailing columns (and the
// right-hand side) with panel k; ONE grid barrier per panel. The panel factorisation needs one block barrier
// per column: a single pass over the column accumulates |tail|^2 AND the dot products with the remaining
// panel columns (they do not depend on beta), every thread then derives beta / tau redundantly, and the
// second pass scales the reflector and updates the remaining columns on rows the thread owns.
// r1 v4 (panel factored between two grid barriers, five block barriers and a one-thread scalar step per
// column): 41 us per 8-column panel.
template <class T>
struct QrSmem {
  // dynamic layout: sV[QR_PB][LV] reflectors of the current panel | sP[QR_PB][LV] panel being factored (CTA 0) |
  // stau[QR_PB] | spart[2][QR_THREADS/32][QR_PB + 1]
  static size_t bytes(int kd) { return ((size_t)2 * QR_PB * (kd + QR_PB) + QR_PB + 2 * (QR_THREADS / 32) * (QR_PB + 1)) * sizeof(T); }
};

// apply the pb reflectors staged in sV (v(col) = 1 implicit, sV[c][r] valid for r > c) to one column / the rhs
template <class T>
__device__ __forceinline__ void qr_apply_panel(const QRMat<T>& Q, const T* __restrict__ sV, const T* __restrict__ stau, const int LV, const int k0,
                                               const int pb, const int j, T* __restrict__ xp, const int rlo, const int rlast, const bool is_rhs,
                                               const int lane, T* __restrict__ sdst = nullptr, const int sbase = 0) {
  T x[QR_MAXR];
#pragma unroll
  for (int t = 0; t < QR_MAXR; ++t) { const int i = k0 + lane + 32 * t; x[t] = (i >= rlo && i <= rlast) ? xp[i] : T(0); }
  for (int c = 0; c < pb; ++c) {
    const int col = k0 + c;
    const T tau = stau[c];
    if ((!is_rhs && j > col + Q.ku) || tau == T(0)) continue;
    const T* v = sV + c * LV;
    T d = T(0);
#pragma unroll
    for (int t = 0; t < QR_MAXR; ++t) { const int r = lane + 32 * t; if (r < LV) d += v[r] * x[t]; }  // v[r] = 0 for r <= c and beyond the column
    d = warp_sum(d);
    const T xcol = __shfl_sync(0xffffffffu, x[0], c);  // x(col): local row c (< QR_PB <= 32) lives in lane c, register 0
    d = (d + xcol) * tau;
#pragma unroll
    for (int t = 0; t < QR_MAXR; ++t) { const int r = lane + 32 * t; if (r < LV) x[t] -= d * v[r]; }
    if (lane == c) x[0] -= d;
  }
  // rows >= sbase of a next-panel column go to the panel CTA's shared-memory panel (they are rewritten by the panel
  // write-back anyway); the rows above are final entries of R
#pragma unroll
  for (int t = 0; t < QR_MAXR; ++t) {
    const int i = k0 + lane + 32 * t;
    if (i >= rlo && i <= rlast) { if (sdst && i >= sbase) sdst[i - sbase] = x[t]; else xp[i] = x[t]; }
  }
}

// panel columns k0 .. k0+pb-1 (rows >= k0) from global memory, factor in shared memory, write back + tau (one CTA)
template <class T>
__device__ __forceinline__ void qr_factor_panel(const QRMat<T>& Q, T* __restrict__ sP, T* __restrict__ spart, T* __restrict__ tauv, const int LV,
                                                const int k0, const int pb, const int rows_staged) {
  // rows_staged: local rows [0, rows_staged) of every column are already in sP (written by qr_apply_panel); the
  // remaining band rows (at most pb per column) come from global memory, everything else is zero
  constexpr int NW = QR_THREADS / 32;
  const int n = Q.n, kd = Q.kd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T tiny = (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f));
  if (rows_staged == 0) {
    for (int idx0 = 0; idx0 < pb * LV; idx0 += 4 * QR_THREADS) {
      T tmp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + tid + QR_THREADS * u, c = idx / LV, r = idx - c * LV, i = k0 + r, j = k0 + c;
        tmp[u] = (idx < pb * LV && i < n && i <= j + kd) ? gq(Q, i, j) : T(0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int idx = idx0 + tid + QR_THREADS * u; if (idx < pb * LV) sP[idx] = tmp[u]; }
    }
  } else {
    const int nrest = LV - rows_staged;  // <= QR_PB local rows per column
    for (int idx = tid; idx < pb * nrest; idx += QR_THREADS) {
      const int c = idx / nrest, r = rows_staged + (idx - c * nrest), i = k0 + r, j = k0 + c;
      sP[c * LV + r] = (i < n && i <= j + kd) ? gq(Q, i, j) : T(0);
    }
  }
  __syncthreads();
  T pend[QR_PB + 1];   // thread 0: row-c entries of column c's step, written after the next barrier
  int pend_c = -1;
  T mytau = T(0);      // thread c keeps tau_c
  for (int c = 0; c < pb; ++c) {
    T* v = sP + c * LV;
    const int rend = min(n - 1, k0 + c + kd) - k0;  // last local row of this column
    T acc[QR_PB];
#pragma unroll
    for (int q = 0; q < QR_PB; ++q) acc[q] = T(0);
    T vr[QR_RPT];
#pragma unroll
    for (int u = 0; u < QR_RPT; ++u) {
      const int r = tid + QR_THREADS * u;
      const bool ok = r > c && r <= rend;
      vr[u] = ok ? v[r] : T(0);
      acc[0] += vr[u] * vr[u];
#pragma unroll
      for (int q = 1; q < QR_PB; ++q) if (c + q < pb && ok) acc[q] += vr[u] * sP[(c + q) * LV + r];
    }
#pragma unroll
    for (int q = 0; q < QR_PB; ++q) acc[q] = warp_sum(acc[q]);
    T* part = spart + (c & 1) * NW * (QR_PB + 1);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < QR_PB; ++q) part[warp * (QR_PB + 1) + q] = acc[q];
    }
    __syncthreads();
    if (tid == 0 && pend_c >= 0) {  // deferred row writes of the previous column (all its readers are past the barrier)
      T* pv = sP + pend_c * LV;
      pv[pend_c] = pend[0];
#pragma unroll
      for (int q = 1; q < QR_PB; ++q) if (pend_c + q < pb) sP[(pend_c + q) * LV + pend_c] = pend[q];
    }
    T tot[QR_PB];
#pragma unroll
    for (int q = 0; q < QR_PB; ++q) { T t2 = T(0); for (int w = 0; w < NW; ++w) t2 += part[w * (QR_PB + 1) + q]; tot[q] = t2; }
    const T c0 = v[c];
    T tau = T(0), inv = T(0), beta = c0;
    if (tot[0] > tiny) {
      beta = sqrt(c0 * c0 + tot[0]);
      if (c0 >= T(0)) beta = -beta;
      inv = T(1) / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    if (tid == c) mytau = tau;
    // t_q = tau * (x_q(c) + inv * <tail, x_q>): x_q -= t_q * v_normalised, v_normalised = inv * tail below row c, 1 at row c
    T tq[QR_PB];
#pragma unroll
    for (int q = 1; q < QR_PB; ++q) tq[q] = (c + q < pb) ? tau * (sP[(c + q) * LV + c] + inv * tot[q]) : T(0);
    if (tid == 0) {
      pend_c = c; pend[0] = beta;
#pragma unroll
      for (int q = 1; q < QR_PB; ++q) if (c + q < pb) pend[q] = sP[(c + q) * LV + c] - tq[q];
    }
    if (tau != T(0)) {
#pragma unroll
      for (int u = 0; u < QR_RPT; ++u) {
        const int r = tid + QR_THREADS * u;
        if (r > c && r <= rend) {
          const T vn = vr[u] * inv;
          v[r] = vn;
#pragma unroll
          for (int q = 1; q < QR_PB; ++q) if (c + q < pb) sP[(c + q) * LV + r] -= tq[q] * vn;
        }
      }
    }
    // no barrier here: the next column's first pass touches only rows this thread owns; the row-c entries read
    // above are rewritten by thread 0 after the next barrier
  }
  __syncthreads();
  if (tid == 0 && pend_c >= 0) {
    T* pv = sP + pend_c * LV;
    pv[pend_c] = pend[0];
#pragma unroll
    for (int q = 1; q < QR_PB; ++q) if (pend_c + q < pb) sP[(pend_c + q) * LV + pend_c] = pend[q];
  }
  if (tid < pb) tauv[k0 + tid] = mytau;
  __syncthreads();
  for (int idx = tid; idx < pb * LV; idx += QR_THREADS) {
    const int c = idx / LV, r = idx - c * LV, i = k0 + r, j = k0 + c;
    if (i < n && i <= j + kd) gq(Q, i, j) = sP[idx];
  }
}

template <class T>
__global__ void __launch_bounds__(QR_THREADS) k_band_qr_reg(QRMat<T> Q, T* __restrict__ tauv, T* __restrict__ rhs, long long* __restrict__ dbg) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char qr_smem_raw[];
#ifdef BA_QR_TICKS
  long long tk_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq_ = clock64();
#define QTICK(i) { if (threadIdx.x == 0) { const long long t1_ = clock64(); tk_[i] += t1_ - tq_; tq_ = t1_; } __syncwarp(); }
#else
#define QTICK(i) {}
#endif
  constexpr int NW = QR_THREADS / 32;
  const int n = Q.n, kd = Q.kd, ku = Q.ku;
  const int LV = kd + QR_PB;                       // rows k0 .. k0+LV-1 cover every reflector of a panel
  T* sV = reinterpret_cast<T*>(qr_smem_raw);       // [QR_PB][LV] reflector vectors of the current panel
  T* sP = sV + QR_PB * LV;                         // [QR_PB][LV] panel being factored (CTA 0)
  T* stau = sV + 2 * QR_PB * LV;                   // [QR_PB]
  T* spart = stau + QR_PB;                         // [2][NW][QR_PB + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool panel_cta = blockIdx.x == 0;
  const int nupd = max(1, (int)gridDim.x - 1);     // CTAs that update trailing columns (all of them when the grid is one CTA)
  const int ucta = (gridDim.x > 1) ? (int)blockIdx.x - 1 : 0;
  if (panel_cta) qr_factor_panel<T>(Q, sP, spart, tauv, LV, 0, min(QR_PB, n), 0);
  __threadfence();
  grid.sync();
  for (int k0 = 0; k0 < n; k0 += QR_PB) {
    const int pb = min(QR_PB, n - k0);
    QTICK(0)
    if (panel_cta) {
      // the panel CTA factored this panel in sP: it becomes sV by swapping the buffers (entries on and above the
      // diagonal are masked instead of re-reading the panel from global memory)
      T* tswap = sV; sV = sP; sP = tswap;
      for (int idx = tid; idx < pb * (QR_PB + 1); idx += QR_THREADS) { const int c = idx / (QR_PB + 1), r = idx - c * (QR_PB + 1); if (r <= c) sV[c * LV + r] = T(0); }
    } else {
      // reflector vectors of this panel -> shared memory (v(col) = 1 implicit; sV[c][r] valid for r > c), loads batched
      for (int idx0 = 0; idx0 < pb * LV; idx0 += 4 * QR_THREADS) {
        T tmp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = idx0 + tid + QR_THREADS * u, c = idx / LV, r = idx - c * LV, i = k0 + r, j = k0 + c;
          tmp[u] = (idx < pb * LV && r > c && i < n && i <= j + kd) ? gq(Q, i, j) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int idx = idx0 + tid + QR_THREADS * u; if (idx < pb * LV) sV[idx] = tmp[u]; }
      }
    }
    if (tid < pb) stau[tid] = tauv[k0 + tid];
    __syncthreads();
    QTICK(1)
    const int jlast = min(n - 1, k0 + pb - 1 + ku);
    const int rlast = min(n - 1, k0 + pb - 1 + kd);  // last row any reflector of the panel touches
    const int k1 = k0 + pb, pb1 = min(QR_PB, n - k1);  // next panel (pb1 <= 0: none)
    if (panel_cta) {
      // next-panel columns: rows [k1, rlast] land in sP (local rows [0, rlast - k1]); zero the staged part first for
      // the rows a column's band does not reach
      const int staged = max(0, rlast - k1 + 1);
      for (int w = warp; w < pb1; w += NW) {
        const int j = k1 + w;   // j <= jlast always (pb1 <= QR_PB <= ku)
        qr_apply_panel<T>(Q, sV, stau, LV, k0, pb, j, Q.G + (size_t)j * Q.ld + (ku - j), max(k0, j - ku), rlast, false, lane, sP + w * LV, k1);
      }
      __syncthreads();
      QTICK(2)
      if (pb1 > 0) qr_factor_panel<T>(Q, sP, spart, tauv, LV, k1, pb1, staged);
      QTICK(3)
    }
    if (!panel_cta || gridDim.x == 1) {
      const int jfirst = k1 + max(pb1, 0);
      const int ntrail = jlast - jfirst + 1;  // may be <= 0
      for (int w = ucta * NW + warp; w < max(ntrail, 0) + 1; w += nupd * NW) {
        const bool is_rhs = (w == max(ntrail, 0));
        const int j = is_rhs ? n : (jfirst + w);
        T* xp = is_rhs ? rhs : (Q.G + (size_t)j * Q.ld + (ku - j));  // x(r) = xp[r]
        qr_apply_panel<T>(Q, sV, stau, LV, k0, pb, j, xp, is_rhs ? k0 : max(k0, j - ku), rlast, is_rhs, lane);
      }
    }
    QTICK(4)
    __threadfence();
    grid.sync();
    QTICK(5)
  }
#ifdef BA_QR_TICKS
  if (dbg && tid == 0 && blockIdx.x <= 1) for (int i = 0; i < 8; ++i) dbg[8 * blockIdx.x + i] = tk_[i];
#endif
#undef QTICK
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

// ---------------------------------------------------------------------------------------------
// The same back substitution on one thread-block cluster. Per 32-column block k (last to first):
//   CTA 0, warp 0: b = rhs_k - R(k, k+1) y_{k+1} (its own look-ahead update), y_k = R_kk^-1 b;
//   every other warp of the cluster: rhs_i -= R(i, k+1) y_{k+1} for the row blocks i < k inside the band
//     (the update of the PREVIOUS block's solution, one step behind, so it never waits for the solve);
//   one cluster barrier.
// Row block i has received every update y_j, j >= i+2, by the end of iteration i+1, and y_{i+1}'s from CTA 0.
// r1 v4 (single CTA, update in the loop): 26 us per block, 13 ms at the synthetic scale.
// ---------------------------------------------------------------------------------------------
constexpr int QRS_THREADS = 256;
template <class T>
__global__ void __launch_bounds__(QRS_THREADS, 1) k_band_qr_backsolve_cluster(QRMat<T> Q, T* __restrict__ rhs, T* __restrict__ y, T sign) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NW = QRS_THREADS / 32;
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = Q.n, ku = Q.ku, nt = (n + NB - 1) / NB;
  const T* const zp = reinterpret_cast<const T*>(ba_zero_word);
  // helper warps: every warp of the cluster except CTA 0's warp 0
  const int hw = rank * NW + warp - 1, nhw = C * NW - 1;
  T yprev = T(0);  // CTA 0 warp 0: y_{k+1}(lane)
  for (int k = nt - 1; k >= 0; --k) {
    const int k0 = k * NB, kn0 = k0 + NB;   // kn0: first column of block k+1
    const bool have_next = (k + 1 < nt);
    if (rank == 0 && warp == 0) {
      // R_kk rows as columns: lane = row; R(k0+lane, k0+c) for c >= lane; R(k0+lane, kn0+c) for the look-ahead update
      T rkk[NB], rkn[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const int gi = k0 + lane, gj = k0 + c, gj2 = kn0 + c;
        const bool ok = gj < n && gi <= gj && gj - gi <= ku && gi < n;
        rkk[c] = *(ok ? &gq(Q, gi, gj) : zp);
        const bool ok2 = have_next && gj2 < n && gi < n && gj2 - gi <= ku;
        rkn[c] = *(ok2 ? &gq(Q, gi, gj2) : zp);
      }
      T b = (k0 + lane < n) ? rhs[k0 + lane] : T(0);
      T u0 = T(0), u1 = T(0);
#pragma unroll
      for (int c = 0; c < NB; c += 2) {
        u0 += rkn[c] * __shfl_sync(FULL, yprev, c);
        u1 += rkn[c + 1] * __shfl_sync(FULL, yprev, c + 1);
      }
      b -= (u0 + u1);
      // diagonal reciprocal (rows past n: identity)
      T dg = T(1);
#pragma unroll
      for (int c = 0; c < NB; ++c) if (c == lane) dg = rkk[c];
      if (k0 + lane >= n) dg = T(1);
      const T inv = T(1) / dg;
#pragma unroll
      for (int j = NB - 1; j >= 0; --j) {
        const T yj = __shfl_sync(FULL, b * inv, j);
        if (lane == j) b = yj;
        if (lane < j) b -= rkk[j] * yj;
      }
      yprev = (k0 + lane < n) ? b : T(0);
      if (k0 + lane < n) y[k0 + lane] = sign * b;
    } else if (have_next) {
      // update with y_{k+1} (published before the previous barrier): row blocks i < k within ku of block k+1's columns
      const T yv = (kn0 + lane < n) ? sign * y[kn0 + lane] : T(0);   // y holds sign * solution; sign = +-1
      const int ilo = max(0, kn0 - ku);                               // first row any column of block k+1 touches
      const int nrows = k0 - ilo;                                     // rows [ilo, k0)
      for (int rb = hw; rb * NB < nrows; rb += nhw) {
        const int gi = ilo + rb * NB + lane;
        const bool rowok = gi < k0;
        T acc0 = T(0), acc1 = T(0);
#pragma unroll 8
        for (int c = 0; c < NB; c += 2) {
          const int gj = kn0 + c;
          const bool ok0 = rowok && gj < n && gj - gi <= ku, ok1 = rowok && gj + 1 < n && gj + 1 - gi <= ku;
          const T r0 = *(ok0 ? &gq(Q, gi, gj) : zp), r1 = *(ok1 ? &gq(Q, gi, gj + 1) : zp);
          acc0 += r0 * __shfl_sync(FULL, yv, c);
          acc1 += r1 * __shfl_sync(FULL, yv, c + 1);
        }
        if (rowok) rhs[gi] -= (acc0 + acc1);
      }
    }
    cluster.sync();
  }
}

}  // namespace ba

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
static_assert(QR_PB == 8 && QR_THREADS == 256, "the reductions of the panel kernel are written for 8-column panels and 8 warps");

// Look-ahead + compact WY. CTA 0 is the PANEL CTA: in iteration k it gives the columns of panel k+1 panel k's
// update and factors panel k+1 in shared memory while every other CTA updates the remaining trailing columns
// (and the right-hand side) with panel k; ONE grid barrier per panel.
//  * Panel factorisation, one block barrier per column: a single pass over the column accumulates |tail|^2, the
//    dot products with the remaining panel columns (they do not depend on beta) and the dot products with the
//    previous reflectors (for T); every thread then derives beta / tau redundantly; the second pass scales the
//    reflector and updates the remaining columns on rows the thread owns.
//  * The panel's reflectors are applied as Q^T x = x - V (T^T (V^T x)) (T: 8x8 upper triangular, dlarft
//    recurrence): the 8 dot products of a column are independent, so the warp does ONE reduction per column
//    instead of eight dependent ones, and one warp carries two columns so that V is read once for both.
// r1 v4 (panel factored between two grid barriers, five block barriers and a one-thread scalar step per column,
// reflectors applied one by one): 41 us per 8-column panel; r1 v6 with look-ahead only: 24.7 us
// (profiles/r01_dense_notes.md: apply 6.4 us, panel 16.3 us).
constexpr int QR_WS = 16 * 33 + 16;   // per-warp scratch of qr_apply_wy: partial dots [16][33] + z [16]
template <class T>
struct QrSmem {
  // dynamic layout: sV[QR_PB][LV] | sP[QR_PB][LV] | sT[64] | sTn[64] | spart[2][64] | sRow[2][8] | sG[64] | ws[8][QR_WS]
  // (the panel CTA's qr_apply_cta uses the ws area as [8][32] partial sums + [64] W / Z)
  static size_t bytes(int kd) { return ((size_t)2 * QR_PB * (kd + QR_PB) + 5 * 64 + 16 + (QR_THREADS / 32) * QR_WS) * sizeof(T); }
};

// Reduce-scatter over the warp followed by a butterfly on the surviving entry: on return every lane holds the warp
// total of entry (lane >> (5 - LOG)) & (NV - 1); NV - 1 + (5 - LOG) shuffles instead of 5 NV.
template <class T, int NV, int LOG>
__device__ __forceinline__ T warp_reduce_scatter(T (&v)[NV], const int lane) {
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int st = 0; st < LOG; ++st) {
    const int h = NV >> (st + 1), off = 16 >> st;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) {
      if (i < h) {
        const T send = upper ? v[i] : v[i + h];
        const T recv = __shfl_xor_sync(FULL, send, off);
        v[i] = (upper ? v[i + h] : v[i]) + recv;
      }
    }
  }
#pragma unroll
  for (int off = 16 >> LOG; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(FULL, v[0], off);
  return v[0];
}

// Householder scalars from the leading entry c0 and |tail|^2 (valid: tail2 > tiny): beta = -sign(c0) |x|,
// inv = 1 / (c0 - beta), tau = (beta - c0) / beta = 1 - c0 / beta. Double: rsqrt + reciprocal seeds with
// FMA corrections (a dependent chain of ~100 cycles instead of the ~400 of sqrt and two IEEE divisions in
// sequence; relative error ~1e-16, which perturbs the reflector within its own rounding error).
__device__ __forceinline__ void householder_scalars(const double c0, const double tail2, double& beta, double& inv, double& tau) {
  const double a = fma(c0, c0, tail2);
  double r = rsqrt(a);                       // 1 / |x|
  const double nb0 = a * r;
  const double nb = fma(fma(-nb0, nb0, a), 0.5 * r, nb0);  // one Newton step on sqrt(a)
  r = fma(fma(-nb, r, 1.0), r, r);           // and on its reciprocal
  beta = (c0 >= 0.0) ? -nb : nb;
  const double rb = (c0 >= 0.0) ? -r : r;    // 1 / beta
  inv = pivot_rcp(c0 - beta);
  tau = fma(-c0, rb, 1.0);
}
__device__ __forceinline__ void householder_scalars(const float c0, const float tail2, float& beta, float& inv, float& tau) {
  beta = sqrtf(c0 * c0 + tail2);
  if (c0 >= 0.0f) beta = -beta;
  inv = 1.0f / (c0 - beta);
  tau = (beta - c0) / beta;
}

// Apply the panel staged in sV (unit diagonal, zeros above; zero beyond each column's band) and sT to two columns
// (or one: xp1 == nullptr). x(r) = xp[r]; rows [rlo, rlast]; rows >= sbase go to sdst (panel CTA) instead of xp.
// The loops over the 8 reflectors are ROLLED (2 KB of code; fully unrolled the function was 44 KB and the kernel
// instruction-fetch bound): pass 1 leaves the per-lane partial dot products in the warp's scratch, 16 lanes sum
// them, z = T^T w by shuffles, pass 2 subtracts V z.
template <class T>
__device__ __noinline__ void qr_apply_wy(const T* __restrict__ sV, const T* __restrict__ sT, T* __restrict__ ws, const int LV, const int k0,
                                         const int rlast, T* __restrict__ xp0, const int rlo0, T* __restrict__ xp1, const int rlo1, const int lane,
                                         T* __restrict__ sdst0, T* __restrict__ sdst1, const int sbase) {
  constexpr unsigned FULL = 0xffffffffu;
  const bool two = xp1 != nullptr;
  T x0[QR_MAXR], x1[QR_MAXR];
#pragma unroll
  for (int t = 0; t < QR_MAXR; ++t) {
    const int i = k0 + lane + 32 * t;
    x0[t] = (i >= rlo0 && i <= rlast) ? xp0[i] : T(0);
    x1[t] = (two && i >= rlo1 && i <= rlast) ? xp1[i] : T(0);
  }
  int roff[QR_MAXR];   // row offsets clamped into the column (x is zero there)
#pragma unroll
  for (int t = 0; t < QR_MAXR; ++t) roff[t] = min(lane + 32 * t, LV - 1);
#pragma unroll 2
  for (int c = 0; c < QR_PB; ++c) {
    const T* vc = sV + c * LV;
    T d0a = T(0), d0b = T(0), d1a = T(0), d1b = T(0);
#pragma unroll
    for (int t = 0; t < QR_MAXR; t += 2) {
      const T va = vc[roff[t]], vb = vc[roff[t + 1]];
      d0a += va * x0[t]; d0b += vb * x0[t + 1];
      d1a += va * x1[t]; d1b += vb * x1[t + 1];
    }
    ws[c * 33 + lane] = d0a + d0b;
    ws[(8 + c) * 33 + lane] = d1a + d1b;
  }
  __syncwarp();
  T tot = T(0);
  {
    const T* wr = ws + (lane & 15) * 33;
    T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
    for (int l = 0; l < 32; l += 4) { s0 += wr[l]; s1 += wr[l + 1]; s2 += wr[l + 2]; s3 += wr[l + 3]; }
    tot = (s0 + s1) + (s2 + s3);   // lanes 0..7: w of column 0, lanes 8..15: w of column 1 (16..31 mirror them)
  }
  // z_j = sum_{i <= j} T(i, j) w_i within the lane's group of 8
  {
    const int j = lane & 7, grp = lane & 8;
    T z = T(0);
#pragma unroll
    for (int i = 0; i < QR_PB; ++i) {
      const T wi = __shfl_sync(FULL, tot, grp + i);
      if (i <= j) z += sT[i * QR_PB + j] * wi;
    }
    __syncwarp();
    if (lane < 16) ws[16 * 33 + lane] = z;
  }
  __syncwarp();
#pragma unroll 2
  for (int c = 0; c < QR_PB; ++c) {
    const T* vc = sV + c * LV;
    const T z0 = ws[16 * 33 + c], z1 = ws[16 * 33 + 8 + c];
#pragma unroll
    for (int t = 0; t < QR_MAXR; ++t) {
      const T vv = vc[roff[t]];
      x0[t] -= vv * z0;
      x1[t] -= vv * z1;
    }
  }
#pragma unroll
  for (int t = 0; t < QR_MAXR; ++t) {
    const int i = k0 + lane + 32 * t;
    if (i >= rlo0 && i <= rlast) { if (sdst0 && i >= sbase) sdst0[i - sbase] = x0[t]; else xp0[i] = x0[t]; }
    if (two && i >= rlo1 && i <= rlast) { if (sdst1 && i >= sbase) sdst1[i - sbase] = x1[t]; else xp1[i] = x1[t]; }
  }
  __syncwarp();
}

#ifdef BA_QR_TICKS
__device__ long long qr_ftick[8];
#endif
// ---- panel CTA, register-resident. Thread t owns the local rows t, t+256, t+512 of all 8 columns (24 scalars).
// The panel factorisation in shared memory was bound by shared-memory bandwidth (every column step re-read and
// re-wrote the remaining columns: 480 KB per panel; r1 v6b: 7.5 us of passes per panel), and so was the
// warp-per-column application on this CTA (every warp read all of V twice: 570 KB).

// x -= V (T^T (V^T x)) for the next panel's columns (global columns k1 .. k1+pb1-1), the whole CTA at once:
// rows relative to k0 in registers, V read once from shared memory, the 64 dot products reduced in two rounds of 32.
// Rows [k0, k1) go back to global memory (final rows of R), rows >= k1 to sP in the factorisation's layout.
template <class T>
__device__ __noinline__ void qr_apply_cta(const QRMat<T>& Q, const T* __restrict__ sV, const T* __restrict__ sT, T* __restrict__ sW, T* __restrict__ sP,
                                          const int LV, const int k0, const int k1, const int pb1, const int rlast) {
  constexpr int NW = QR_THREADS / 32;
  const int n = Q.n, ku = Q.ku;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int ru[QR_RPT]; bool rin[QR_RPT];
#pragma unroll
  for (int u = 0; u < QR_RPT; ++u) { const int r = tid + QR_THREADS * u; rin[u] = r < LV; ru[u] = min(r, LV - 1); }
  T x[QR_PB][QR_RPT];
#pragma unroll
  for (int q = 0; q < QR_PB; ++q)
#pragma unroll
    for (int u = 0; u < QR_RPT; ++u) {
      const int i = k0 + ru[u], j = k1 + q;
      const bool ok = rin[u] && q < pb1 && i <= rlast && i >= j - ku && i < n;
      x[q][u] = ok ? Q.G[(size_t)j * Q.ld + (i - j + ku)] : T(0);
    }
  T* part = sW;            // [NW][32]
  T* sWt = sW + NW * 32;   // [64] W, then Z
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    T w[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) w[e] = T(0);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = 4 * half + cc;
#pragma unroll
      for (int u = 0; u < QR_RPT; ++u) {
        const T vv = rin[u] ? sV[c * LV + ru[u]] : T(0);
#pragma unroll
        for (int q = 0; q < QR_PB; ++q) w[cc * 8 + q] += vv * x[q][u];
      }
    }
    const T mine = warp_reduce_scatter<T, 32, 5>(w, lane);  // lane holds entry lane
    part[warp * 32 + lane] = mine;
    __syncthreads();
    if (tid < 32) {
      T sacc = T(0);
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) sacc += part[w2 * 32 + tid];
      sWt[32 * half + tid] = sacc;   // W(c, q) at c * 8 + q
    }
    __syncthreads();
  }
  // Z = T^T W: Z(j, q) = sum_{i <= j} T(i, j) W(i, q)
  T zreg = T(0);
  if (tid < 64) {
    const int j = tid >> 3, q = tid & 7;
#pragma unroll
    for (int i = 0; i < QR_PB; ++i) if (i <= j) zreg += sT[i * QR_PB + j] * sWt[i * 8 + q];
  }
  __syncthreads();
  if (tid < 64) sWt[tid] = zreg;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < QR_PB; ++c) {
    T z[QR_PB];
#pragma unroll
    for (int q = 0; q < QR_PB; ++q) z[q] = sWt[c * 8 + q];
#pragma unroll
    for (int u = 0; u < QR_RPT; ++u) {
      const T vv = rin[u] ? sV[c * LV + ru[u]] : T(0);
#pragma unroll
      for (int q = 0; q < QR_PB; ++q) x[q][u] -= vv * z[q];
    }
  }
#pragma unroll
  for (int q = 0; q < QR_PB; ++q)
#pragma unroll
    for (int u = 0; u < QR_RPT; ++u) {
      const int i = k0 + ru[u], j = k1 + q;
      if (rin[u] && q < pb1 && i <= rlast && i >= j - ku && i < n) {
        if (i >= k1) sP[q * LV + (i - k1)] = x[q][u]; else Q.G[(size_t)j * Q.ld + (i - j + ku)] = x[q][u];
      }
    }
}

// Factor the panel columns k0 .. k0+pb-1 (one CTA), panel in registers: local rows [0, rows_staged) of every column
// come from sP (written by qr_apply_cta), the remaining band rows from global memory (rows_staged == 0: everything
// from global memory). One block barrier per column: a single pass accumulates |tail|^2, the dot products with the
// remaining columns and with the previous reflectors (for T); every thread then derives beta / tau redundantly and
// updates its rows. Leaves the factored panel in sP (it becomes sV), in global memory, tau in tauv, T in sTout / Tg.
template <class T, int RPT>
__device__ __noinline__ void qr_factor_panel(const QRMat<T>& Q, T* __restrict__ sP, T* __restrict__ spart, T* __restrict__ sG, T* __restrict__ sTout,
                                             T* __restrict__ tauv, T* __restrict__ Tg, const int LV, const int k0, const int pb,
                                             const int rows_staged) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NW = QR_THREADS / 32;
  const int n = Q.n, kd = Q.kd, ku = Q.ku;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T tiny = (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f));
#ifdef BA_QR_TICKS
  long long ft_ = clock64();
#define FTICK(i) { if (threadIdx.x == 0) { const long long t1_ = clock64(); qr_ftick[i] += t1_ - ft_; ft_ = t1_; } __syncwarp(); }
#else
#define FTICK(i) {}
#endif
  T* sRow = spart + 128;   // [2][8] row c of every column, published by the thread that owns it
  int ru[RPT]; bool rin[RPT];
#pragma unroll
  for (int u = 0; u < RPT; ++u) { const int r = tid + QR_THREADS * u; rin[u] = r < LV; ru[u] = min(r, LV - 1); }
  T p[QR_PB][RPT];
#pragma unroll
  for (int c = 0; c < QR_PB; ++c)
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const int r = ru[u], i = k0 + r, j = k0 + c;
      const bool inband = rin[u] && c < pb && i < n && i <= j + kd;
      T v = T(0);
      if (inband) v = (r < rows_staged) ? sP[c * LV + r] : Q.G[(size_t)j * Q.ld + (r - c + ku)];
      p[c][u] = v;
    }
  if (tid < 64) sG[tid] = T(0);
  T mytau = T(0);      // thread c: tau_c
  FTICK(0)
#pragma unroll
  for (int c = 0; c < QR_PB; ++c) {
    if (c < pb) {
      const int rend = min(n - 1, k0 + c + kd) - k0;  // last local row of this column
      if (tid == c) {
#pragma unroll
        for (int q = 0; q < QR_PB; ++q) sRow[(c & 1) * 8 + q] = p[q][0];
      }
      // slot 0: |tail|^2; slots 1 .. 7-c: <tail, x_{c+s}>; slots 8-c .. 7: <tail, v_i>, i = s-(8-c)
      T acc[QR_PB], vr[RPT];
#pragma unroll
      for (int q = 0; q < QR_PB; ++q) acc[q] = T(0);
#pragma unroll
      for (int u = 0; u < RPT; ++u) {
        const bool ok = rin[u] && ru[u] > c && ru[u] <= rend;
        vr[u] = ok ? p[c][u] : T(0);
        acc[0] += vr[u] * vr[u];
#pragma unroll
        for (int q = 1; q < QR_PB; ++q) acc[q] += vr[u] * p[(q <= 7 - c) ? c + q : q - (8 - c)][u];
      }
      FTICK(1)
      const T mine = warp_reduce_scatter<T, 8, 3>(acc, lane);  // lane holds slot (lane >> 2) & 7
      T* part = spart + (c & 1) * 64;
      if ((lane & 3) == 0) part[warp * 8 + (lane >> 2)] = mine;
      __syncthreads();
      FTICK(2)
      T ssum;
      {
        const T* pp = part + (lane & 7);
        ssum = ((pp[0] + pp[8]) + (pp[16] + pp[24])) + ((pp[32] + pp[40]) + (pp[48] + pp[56]));
      }
      T tot[QR_PB], xrow[QR_PB];
#pragma unroll
      for (int q = 0; q < QR_PB; ++q) { tot[q] = __shfl_sync(FULL, ssum, q); xrow[q] = sRow[(c & 1) * 8 + q]; }
      const T c0 = xrow[c];
      T tau = T(0), inv = T(0), beta = c0;
      if (tot[0] > tiny) householder_scalars(c0, tot[0], beta, inv, tau);
      if (tid == c) mytau = tau;
      FTICK(3)
      // t_q = tau * (x_q(c) + inv * <tail, x_q>): x_q -= t_q * v_normalised (inv * tail below row c, 1 at row c)
      T tq[QR_PB];
#pragma unroll
      for (int q = 0; q < QR_PB; ++q) tq[q] = (q > c) ? tau * (xrow[q] + inv * tot[(q > c) ? q - c : 0]) : T(0);
      // g_i = <v_i, v_c> = v_i(c) + inv * <tail, v_i> for the previous reflectors (T is formed after the loop)
      if (c > 0 && warp == NW - 1 && lane >= 8 - c && lane < 8) {
        T xr = T(0);
#pragma unroll
        for (int i = 0; i < QR_PB; ++i) if (i == lane - (8 - c)) xr = xrow[i];
        sG[c * QR_PB + (lane - (8 - c))] = xr + inv * ssum;
      }
      FTICK(4)
#pragma unroll
      for (int u = 0; u < RPT; ++u) {
        if (rin[u] && ru[u] > c && ru[u] <= rend) {
          const T vn = vr[u] * inv;   // tau == 0: inv == 0, the stored tail reads as zero (H = I)
          p[c][u] = vn;
#pragma unroll
          for (int q = 0; q < QR_PB; ++q) if (q > c) p[q][u] -= tq[q] * vn;
        }
      }
      if (tid == c) {   // row c of the remaining columns, and the diagonal
        p[c][0] = beta;
#pragma unroll
        for (int q = 0; q < QR_PB; ++q) if (q > c) p[q][0] -= tq[q];
      }
      FTICK(5)
    }
  }
  __syncthreads();
  if (tid < pb) tauv[k0 + tid] = mytau;
  // T (dlarft, forward columnwise): T(c, c) = tau_c, T(0:c, c) = -tau_c T(0:c, 0:c) g(0:c, c); thread i owns row i
  if (tid < QR_PB) {
    T trow[QR_PB];
#pragma unroll
    for (int c = 0; c < QR_PB; ++c) {
      const T tc = __shfl_sync(0xffu, mytau, c);
      T sacc = T(0);
#pragma unroll
      for (int m = 0; m < QR_PB; ++m) if (m < c) sacc += ((m >= tid) ? trow[m] : T(0)) * sG[c * QR_PB + m];
      trow[c] = (tid == c) ? tc : ((tid < c) ? -tc * sacc : T(0));
    }
#pragma unroll
    for (int q = 0; q < QR_PB; ++q) { sTout[tid * QR_PB + q] = trow[q]; Tg[tid * QR_PB + q] = trow[q]; }
  }
  // the factored panel: to sP (next iteration's sV) and to the band storage
#pragma unroll
  for (int c = 0; c < QR_PB; ++c)
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const int r = ru[u], i = k0 + r, j = k0 + c;
      if (sP != nullptr && rin[u]) sP[c * LV + r] = p[c][u];
      if (rin[u] && c < pb && i < n && i <= j + kd) Q.G[(size_t)j * Q.ld + (r - c + ku)] = p[c][u];
    }
  FTICK(6)
#undef FTICK
}

template <class T>
__global__ void __launch_bounds__(QR_THREADS) k_band_qr_reg(QRMat<T> Q, T* __restrict__ tauv, T* __restrict__ rhs, T* __restrict__ Tg2,
                                                            long long* __restrict__ dbg) {
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
  T* sT = sV + 2 * QR_PB * LV;                     // [64] T of the current panel
  T* sTn = sT + 64;                                // [64] T of the panel being factored (CTA 0)
  T* spart = sTn + 64;                             // [2][8][8]
  T* sG = spart + 128 + 16;                        // [64] (spart is followed by sRow[2][8])
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* ws = sG + 64 + warp * QR_WS;                  // per-warp scratch of qr_apply_wy
  const bool panel_cta = blockIdx.x == 0;
  const int nupd = max(1, (int)gridDim.x - 1);     // CTAs that update trailing columns (all of them when the grid is one CTA)
  const int ucta = (gridDim.x > 1) ? (int)blockIdx.x - 1 : 0;
  if (panel_cta) qr_factor_panel<T, QR_RPT>(Q, sP, spart, sG, sTn, tauv, Tg2, LV, 0, min(QR_PB, n), 0);
  __threadfence();
  grid.sync();
  int par = 0;  // parity of the current panel: its T sits in Tg2 + 64 * par
  for (int k0 = 0; k0 < n; k0 += QR_PB, par ^= 1) {
    const int pb = min(QR_PB, n - k0);
    QTICK(0)
    if (panel_cta) {
      // the panel CTA factored this panel in sP / sTn: they become sV / sT by swapping the buffers
      T* tswap = sV; sV = sP; sP = tswap;
      tswap = sT; sT = sTn; sTn = tswap;
      for (int idx = tid; idx < QR_PB * (QR_PB + 1); idx += QR_THREADS) {
        const int c = idx / (QR_PB + 1), r = idx - c * (QR_PB + 1);
        if (r <= c) sV[c * LV + r] = (r == c && c < pb) ? T(1) : T(0);
      }
      if (pb < QR_PB) for (int idx = tid; idx < (QR_PB - pb) * LV; idx += QR_THREADS) sV[pb * LV + idx] = T(0);
    } else {
      // reflector vectors of this panel -> shared memory (unit diagonal, zeros above): 24 independent loads per thread
      T tmp[QR_PB][QR_RPT];
#pragma unroll
      for (int c = 0; c < QR_PB; ++c)
#pragma unroll
        for (int u = 0; u < QR_RPT; ++u) {
          const int r = tid + QR_THREADS * u, i = k0 + r, j = k0 + c;
          const bool ld_ = r < LV && c < pb && r > c && i < n && i <= j + kd;
          tmp[c][u] = ld_ ? Q.G[(size_t)j * Q.ld + (r - c + ku)] : ((r == c && c < pb) ? T(1) : T(0));
        }
#pragma unroll
      for (int c = 0; c < QR_PB; ++c)
#pragma unroll
        for (int u = 0; u < QR_RPT; ++u) { const int r = tid + QR_THREADS * u; if (r < LV) sV[c * LV + r] = tmp[c][u]; }
      if (tid < 64) sT[tid] = Tg2[64 * par + tid];
    }
    __syncthreads();
    QTICK(1)
    const int jlast = min(n - 1, k0 + pb - 1 + ku);
    const int rlast = min(n - 1, k0 + pb - 1 + kd);  // last row any reflector of the panel touches
    const int k1 = k0 + pb, pb1 = min(QR_PB, n - k1);  // next panel (pb1 <= 0: none)
    if (panel_cta) {
      // next-panel columns, the whole CTA at once: rows [k1, rlast] land in sP (local rows [0, rlast - k1])
      const int staged = max(0, rlast - k1 + 1);
      if (pb1 > 0) qr_apply_cta<T>(Q, sV, sT, sG + 64, sP, LV, k0, k1, pb1, rlast);
      __syncthreads();
      QTICK(2)
      if (pb1 > 0) qr_factor_panel<T, QR_RPT>(Q, sP, spart, sG, sTn, tauv, Tg2 + 64 * (par ^ 1), LV, k1, pb1, staged);
      QTICK(3)
    }
    if (!panel_cta || gridDim.x == 1) {
      const int jfirst = k1 + max(pb1, 0);
      const int ntrail = max(0, jlast - jfirst + 1);  // trailing columns; item ntrail is the right-hand side
      const int nitems = ntrail + 1, ntask = (nitems + 1) / 2;
      for (int w = ucta * NW + warp; w < ntask; w += nupd * NW) {
        const int it0 = 2 * w, it1 = 2 * w + 1;
        const bool rhs0 = it0 == ntrail, has1 = it1 < nitems, rhs1 = it1 == ntrail;
        const int j0 = jfirst + it0, j1 = jfirst + it1;
        T* xp0 = rhs0 ? rhs : (Q.G + (size_t)j0 * Q.ld + (ku - j0));
        T* xp1 = !has1 ? nullptr : (rhs1 ? rhs : (Q.G + (size_t)j1 * Q.ld + (ku - j1)));
        qr_apply_wy<T>(sV, sT, ws, LV, k0, rlast, xp0, rhs0 ? k0 : max(k0, j0 - ku), xp1, rhs1 ? k0 : max(k0, j1 - ku), lane, nullptr, nullptr, 0);
      }
    }
    QTICK(4)
    __threadfence();
    grid.sync();
    QTICK(5)
  }
#ifdef BA_QR_TICKS
  if (dbg && tid == 0 && blockIdx.x == 0) for (int i = 0; i < 8; ++i) { dbg[i] = tk_[i]; dbg[8 + i] = qr_ftick[i]; qr_ftick[i] = 0; }
#endif
#undef QTICK
}

// ---------------------------------------------------------------------------------------------
// Tall variant for reduced systems whose columns do not fit the register-resident update (kd + 8 > 640: dense S of
// the bundled problems with more than 70 cameras). Same structure (panel CTA with look-ahead, compact WY, one grid
// barrier per panel), but a trailing column is STREAMED twice from L2 in row chunks (w = V^T x, then x -= V z) instead
// of living in registers, the panel CTA keeps no second panel buffer (the factored panel goes to global memory and
// is staged like on every other CTA), and the panel factorisation holds RPT rows per thread. The previous
// fallback (k_band_qr: reflectors applied one by one from global memory, 8 x the L2 traffic, panel factored in
// global memory between two grid barriers) took 34 ms for 126 cameras and 115 ms for 257.
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __noinline__ void qr_apply_wy_stream(const T* __restrict__ sV, const T* __restrict__ sT, const int LV, const int k0, const int rlast,
                                                T* __restrict__ xp0, const int rlo0, T* __restrict__ xp1, const int rlo1, const int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  const bool two = xp1 != nullptr;
  const int nrow = rlast - k0 + 1;   // local rows [0, nrow)
  T w[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) w[e] = T(0);
  for (int r = lane; r < nrow; r += 32) {
    const int i = k0 + r;
    const T x0 = (i >= rlo0) ? xp0[i] : T(0);
    const T x1 = (two && i >= rlo1) ? xp1[i] : T(0);
#pragma unroll
    for (int c = 0; c < QR_PB; ++c) { const T vv = sV[c * LV + r]; w[c] += vv * x0; w[8 + c] += vv * x1; }
  }
  const T mine = warp_reduce_scatter<T, 16, 4>(w, lane);  // lane holds entry (lane >> 1) & 15
  T z0[QR_PB], z1[QR_PB];
  {
    T wt[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) wt[e] = __shfl_sync(FULL, mine, 2 * e);
#pragma unroll
    for (int j = 0; j < QR_PB; ++j) {
      T a0 = T(0), a1 = T(0);
#pragma unroll
      for (int i = 0; i < QR_PB; ++i) if (i <= j) { const T tij = sT[i * QR_PB + j]; a0 += tij * wt[i]; a1 += tij * wt[8 + i]; }
      z0[j] = a0; z1[j] = a1;
    }
  }
  for (int r = lane; r < nrow; r += 32) {
    const int i = k0 + r;
    T x0 = (i >= rlo0) ? xp0[i] : T(0);
    T x1 = (two && i >= rlo1) ? xp1[i] : T(0);
#pragma unroll
    for (int c = 0; c < QR_PB; ++c) { const T vv = sV[c * LV + r]; x0 -= vv * z0[c]; x1 -= vv * z1[c]; }
    if (i >= rlo0) xp0[i] = x0;
    if (two && i >= rlo1) xp1[i] = x1;
  }
}

template <class T>
struct QrTallSmem {
  // dynamic layout: sV[QR_PB][LV] | sT[64] | sTn[64] | spart[2][64] | sRow[2][8] | sG[64]
  static size_t bytes(int kd) { return ((size_t)QR_PB * (kd + QR_PB) + 5 * 64 + 16) * sizeof(T); }
};

template <class T, int RPT>
__global__ void __launch_bounds__(QR_THREADS) k_band_qr_tall(QRMat<T> Q, T* __restrict__ tauv, T* __restrict__ rhs, T* __restrict__ Tg2) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char qr_smem_raw[];
  constexpr int NW = QR_THREADS / 32;
  const int n = Q.n, kd = Q.kd, ku = Q.ku;
  const int LV = kd + QR_PB;
  T* sV = reinterpret_cast<T*>(qr_smem_raw);
  T* sT = sV + QR_PB * LV;
  T* sTn = sT + 64;
  T* spart = sTn + 64;       // [2][64], followed by sRow[2][8]
  T* sG = spart + 128 + 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool panel_cta = blockIdx.x == 0;
  const int nupd = max(1, (int)gridDim.x - 1);
  const int ucta = (gridDim.x > 1) ? (int)blockIdx.x - 1 : 0;
  if (panel_cta) qr_factor_panel<T, RPT>(Q, nullptr, spart, sG, sTn, tauv, Tg2, LV, 0, min(QR_PB, n), 0);
  __threadfence();
  grid.sync();
  int par = 0;
  for (int k0 = 0; k0 < n; k0 += QR_PB, par ^= 1) {
    const int pb = min(QR_PB, n - k0);
    // reflector vectors of this panel -> shared memory (unit diagonal, zeros above and beyond the band), loads batched
    for (int idx0 = 0; idx0 < QR_PB * LV; idx0 += 8 * QR_THREADS) {
      T tmp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = idx0 + tid + QR_THREADS * u, c = idx / LV, r = idx - c * LV, i = k0 + r, j = k0 + c;
        const bool ld_ = idx < QR_PB * LV && c < pb && r > c && i < n && i <= j + kd;
        tmp[u] = ld_ ? Q.G[(size_t)j * Q.ld + (r - c + ku)] : ((idx < QR_PB * LV && r == c && c < pb) ? T(1) : T(0));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int idx = idx0 + tid + QR_THREADS * u; if (idx < QR_PB * LV) sV[idx] = tmp[u]; }
    }
    if (tid < 64) sT[tid] = Tg2[64 * par + tid];
    __syncthreads();
    const int jlast = min(n - 1, k0 + pb - 1 + ku);
    const int rlast = min(n - 1, k0 + pb - 1 + kd);
    const int k1 = k0 + pb, pb1 = min(QR_PB, n - k1);
    if (panel_cta) {
      for (int w = 2 * warp; w < pb1; w += 2 * NW) {
        const int j0 = k1 + w, j1 = j0 + 1;
        const bool two = w + 1 < pb1;
        qr_apply_wy_stream<T>(sV, sT, LV, k0, rlast, Q.G + (size_t)j0 * Q.ld + (ku - j0), max(k0, j0 - ku),
                              two ? Q.G + (size_t)j1 * Q.ld + (ku - j1) : nullptr, max(k0, j1 - ku), lane);
      }
      __syncthreads();
      if (pb1 > 0) qr_factor_panel<T, RPT>(Q, nullptr, spart, sG, sTn, tauv, Tg2 + 64 * (par ^ 1), LV, k1, pb1, 0);
    }
    if (!panel_cta || gridDim.x == 1) {
      const int jfirst = k1 + max(pb1, 0);
      const int ntrail = max(0, jlast - jfirst + 1);
      const int nitems = ntrail + 1, ntask = (nitems + 1) / 2;
      for (int w = ucta * NW + warp; w < ntask; w += nupd * NW) {
        const int it0 = 2 * w, it1 = 2 * w + 1;
        const bool rhs0 = it0 == ntrail, has1 = it1 < nitems, rhs1 = it1 == ntrail;
        const int j0 = jfirst + it0, j1 = jfirst + it1;
        T* xp0 = rhs0 ? rhs : (Q.G + (size_t)j0 * Q.ld + (ku - j0));
        T* xp1 = !has1 ? nullptr : (rhs1 ? rhs : (Q.G + (size_t)j1 * Q.ld + (ku - j1)));
        qr_apply_wy_stream<T>(sV, sT, LV, k0, rlast, xp0, rhs0 ? k0 : max(k0, j0 - ku), xp1, rhs1 ? k0 : max(k0, j1 - ku), lane);
      }
    }
    __threadfence();
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

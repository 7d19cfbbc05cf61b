// Point-tile kernels for sm_100a: the fused Jacobian + per-point block factorisation + Schur
// accumulation kernel (K2) and the fused back-substitution + update + test-energy kernel (K45).
//
// Replaces, on the device and without ever materialising J:
//   row permutation + [J; sqrt(lambda) I]                  BacktrackLevMarqQRChol.h:291-315
//   BlockDiagonalSparseQR<.., ColPivHouseholderQR>::compute on the (2n_j+3)x3 point blocks and
//   Q1^T * J2 (camera columns), matrixQ().transpose()*r     QRChol.h:319-329 (solver NOT IN TREE)
//   J2bot^T J2bot, J2bot^T qtb2                             QRChol.h:339-341
//   back-substitution + colsPermutation                    QRChol.h:344-360
//   increment_in_place + functor(xTest)                    QRChol.h:363-371, BAFunctor.h:299-342
//
// Work decomposition (B200-first): a CTA owns a TILE of consecutive points whose
// observations fit TILE lanes. Phase 1 is one lane per OBSERVATION (full-lane Jacobians), phase 2
// one lane per POINT (3-column Householder QR with column pivoting, thin Q formed in place in
// shared memory), phase 3 one lane per observation (R12_i = Q1_i^T Jc_i), phase 4 one WARP per
// camera pair of a point with lanes <-> the 81 entries of the 9x9 block, so that the atomics into
// the reduced camera matrix are coalesced 72-byte row segments.
#pragma once
#include "ba_model.cuh"

namespace ba {

constexpr int TILE = 128;       // lanes per CTA = max observations (and points) per tile
constexpr int TP = TILE + 1;    // odd SoA stride -> conflict-free both per-lane and per-row

enum PointFactor { PF_HOUSEHOLDER = 0, PF_NORMAL = 1 };

template <class T> struct TileSmem {
  T Q[6 * TP];     // Jp, overwritten by the thin Q1 rows (2x3 per observation)
  T E[2 * TP];     // residual
  T Jc[18 * TP];   // camera Jacobian block (2x9)
  T R12[27 * TP];  // Q1_i^T Jc_i (3x9)
  T Rm[6 * TP];    // per point: r00 r01 r02 r11 r12 r22
  T C[3 * TP];     // per point: c = Q1^T e
  T G[3 * TP];     // per point: Jp^T e (for JtRes) -> reused as u / dx
  int perm[TP];    // per point: packed column permutation
  int ptObs0[TP];  // per point: first local observation
  int ptN[TP];     // per point: observation count
  int pairOff[TP + 1];
  double red[3 * (TILE / 32)];
};

template <class T> __device__ __forceinline__ void atomic_add(T* p, T v) { atomicAdd(p, v); }

// ---------------------------------------------------------------------------------------------
// Phase 2: one lane per point. Rows rho = 2*i + a live in sm.Q[(3a+b)*TP + lo + i]; the three
// lambda rows sqrt(lambda) I3 live in registers. Eigen ColPivHouseholderQR conventions.
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T& qel(TileSmem<T>& sm, int lo, int rho, int b) { return sm.Q[(3 * (rho & 1) + b) * TP + lo + (rho >> 1)]; }

template <class T>
__device__ void point_householder(TileSmem<T>& sm, int p, int lo, int n, T sl) {
  const int m = 2 * n;  // observation rows; plus 3 register rows
  T L[3][3] = {{sl, T(0), T(0)}, {T(0), sl, T(0)}, {T(0), T(0), sl}};
  T tau[3] = {T(0), T(0), T(0)};
  T Rv[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  int pm0 = 0, pm1 = 1, pm2 = 2;
  // g_p = Jp^T e in the ORIGINAL column order
  {
    T g0 = T(0), g1 = T(0), g2 = T(0);
    for (int i = 0; i < n; ++i) {
      const T e0 = sm.E[lo + i], e1 = sm.E[TP + lo + i];
      g0 += sm.Q[0 * TP + lo + i] * e0 + sm.Q[3 * TP + lo + i] * e1;
      g1 += sm.Q[1 * TP + lo + i] * e0 + sm.Q[4 * TP + lo + i] * e1;
      g2 += sm.Q[2 * TP + lo + i] * e0 + sm.Q[5 * TP + lo + i] * e1;
    }
    sm.G[p] = g0; sm.G[TP + p] = g1; sm.G[2 * TP + p] = g2;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // squared norms of the remaining columns over rows >= k (pivot rule)
    T nn[3] = {T(-1), T(-1), T(-1)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < k) continue;
      T s = T(0);
      for (int r = k; r < m; ++r) { const T v = qel(sm, lo, r, c); s += v * v; }
#pragma unroll
      for (int r = 0; r < 3; ++r) s += L[r][c] * L[r][c];
      nn[c] = s;
    }
    int best = k;
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k && nn[c] > nn[best]) best = c;
    if (best != k) {
      for (int r = 0; r < m; ++r) { T& a = qel(sm, lo, r, k); T& b = qel(sm, lo, r, best); const T t = a; a = b; b = t; }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) if (c == best) { const T t = L[r][k]; L[r][k] = L[r][c]; L[r][c] = t; }
      }
      // swap perm entries k <-> best
      int pk = (k == 0) ? pm0 : (k == 1 ? pm1 : pm2);
      int pb = (best == 1) ? pm1 : pm2;
      if (k == 0) pm0 = pb; else if (k == 1) pm1 = pb; else pm2 = pb;
      if (best == 1) pm1 = pk; else pm2 = pk;
    }
    const T c0 = qel(sm, lo, k, k);
    T tail2 = T(0);
    for (int r = k + 1; r < m; ++r) { const T v = qel(sm, lo, r, k); tail2 += v * v; }
#pragma unroll
    for (int r = 0; r < 3; ++r) tail2 += L[r][k] * L[r][k];
    T beta, tk;
    if (tail2 <= (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f))) {
      tk = T(0); beta = c0;
      for (int r = k + 1; r < m; ++r) qel(sm, lo, r, k) = T(0);
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][k] = T(0);
    } else {
      beta = tsqrt(c0 * c0 + tail2);
      if (c0 >= T(0)) beta = -beta;
      const T inv = T(1) / (c0 - beta);
      for (int r = k + 1; r < m; ++r) qel(sm, lo, r, k) *= inv;
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][k] *= inv;
      tk = (beta - c0) / beta;
    }
    tau[k] = tk;
    qel(sm, lo, k, k) = beta;
    // apply H_k to the remaining columns
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c <= k) continue;
      T s = qel(sm, lo, k, c);
      for (int r = k + 1; r < m; ++r) s += qel(sm, lo, r, k) * qel(sm, lo, r, c);
#pragma unroll
      for (int r = 0; r < 3; ++r) s += L[r][k] * L[r][c];
      s *= tk;
      qel(sm, lo, k, c) -= s;
      for (int r = k + 1; r < m; ++r) qel(sm, lo, r, c) -= s * qel(sm, lo, r, k);
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][c] -= s * L[r][k];
    }
  }
  // R (upper triangle) sits in observation rows 0..2 (n >= 2 => m >= 4); column swaps above were
  // applied to all rows, so it is already in pivoted column order.
  Rv[0] = qel(sm, lo, 0, 0); Rv[1] = qel(sm, lo, 0, 1); Rv[2] = qel(sm, lo, 0, 2);
  Rv[3] = qel(sm, lo, 1, 1); Rv[4] = qel(sm, lo, 1, 2); Rv[5] = qel(sm, lo, 2, 2);
#pragma unroll
  for (int i = 0; i < 6; ++i) sm.Rm[i * TP + p] = Rv[i];
  sm.perm[p] = pm0 | (pm1 << 2) | (pm2 << 4);
  // form the thin Q1 in place (dorg2r): k = 2, 1, 0
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const T tk = tau[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c <= k) continue;
      T s = qel(sm, lo, k, c);
      for (int r = k + 1; r < m; ++r) s += qel(sm, lo, r, k) * qel(sm, lo, r, c);
#pragma unroll
      for (int r = 0; r < 3; ++r) s += L[r][k] * L[r][c];
      s *= tk;
      qel(sm, lo, k, c) -= s;
      for (int r = k + 1; r < m; ++r) qel(sm, lo, r, c) -= s * qel(sm, lo, r, k);
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][c] -= s * L[r][k];
    }
    for (int r = k + 1; r < m; ++r) qel(sm, lo, r, k) *= -tk;
#pragma unroll
    for (int r = 0; r < 3; ++r) L[r][k] *= -tk;
    qel(sm, lo, k, k) = T(1) - tk;
    for (int r = 0; r < k; ++r) qel(sm, lo, r, k) = T(0);
  }
  // c = Q1^T e
  T c0 = T(0), c1 = T(0), c2 = T(0);
  for (int i = 0; i < n; ++i) {
    const T e0 = sm.E[lo + i], e1 = sm.E[TP + lo + i];
    c0 += sm.Q[0 * TP + lo + i] * e0 + sm.Q[3 * TP + lo + i] * e1;
    c1 += sm.Q[1 * TP + lo + i] * e0 + sm.Q[4 * TP + lo + i] * e1;
    c2 += sm.Q[2 * TP + lo + i] * e0 + sm.Q[5 * TP + lo + i] * e1;
  }
  sm.C[p] = c0; sm.C[TP + p] = c1; sm.C[2 * TP + p] = c2;
}

// Normal-equation point factor (CHOLESKY variant, BacktrackLevMarqCholesky.h:260-282):
// V = Jp^T Jp + lambda I = L D L^T; R := D^{1/2} L^T; Q1 rows := Jp rows * R^{-1}.
template <class T>
__device__ void point_normal(TileSmem<T>& sm, int p, int lo, int n, T lambda) {
  T v00 = lambda, v10 = T(0), v11 = lambda, v20 = T(0), v21 = T(0), v22 = lambda;
  T g0 = T(0), g1 = T(0), g2 = T(0);
  for (int i = 0; i < n; ++i) {
    const T a0 = sm.Q[0 * TP + lo + i], a1 = sm.Q[1 * TP + lo + i], a2 = sm.Q[2 * TP + lo + i];
    const T b0 = sm.Q[3 * TP + lo + i], b1 = sm.Q[4 * TP + lo + i], b2 = sm.Q[5 * TP + lo + i];
    const T e0 = sm.E[lo + i], e1 = sm.E[TP + lo + i];
    v00 += a0 * a0 + b0 * b0; v10 += a1 * a0 + b1 * b0; v11 += a1 * a1 + b1 * b1;
    v20 += a2 * a0 + b2 * b0; v21 += a2 * a1 + b2 * b1; v22 += a2 * a2 + b2 * b2;
    g0 += a0 * e0 + b0 * e1; g1 += a1 * e0 + b1 * e1; g2 += a2 * e0 + b2 * e1;
  }
  sm.G[p] = g0; sm.G[TP + p] = g1; sm.G[2 * TP + p] = g2;
  const T d0 = v00, l10 = v10 / d0, l20 = v20 / d0;
  const T d1 = v11 - l10 * l10 * d0, l21 = (v21 - l20 * l10 * d0) / d1;
  const T d2 = v22 - l20 * l20 * d0 - l21 * l21 * d1;
  const T s0 = tsqrt(d0), s1 = tsqrt(d1), s2 = tsqrt(d2);
  const T r00 = s0, r01 = s0 * l10, r02 = s0 * l20, r11 = s1, r12 = s1 * l21, r22 = s2;
  sm.Rm[0 * TP + p] = r00; sm.Rm[1 * TP + p] = r01; sm.Rm[2 * TP + p] = r02;
  sm.Rm[3 * TP + p] = r11; sm.Rm[4 * TP + p] = r12; sm.Rm[5 * TP + p] = r22;
  sm.perm[p] = 0 | (1 << 2) | (2 << 4);
  const T i00 = T(1) / r00, i11 = T(1) / r11, i22 = T(1) / r22;
  T c0 = T(0), c1 = T(0), c2 = T(0);
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const T x0 = sm.Q[(3 * a + 0) * TP + lo + i], x1 = sm.Q[(3 * a + 1) * TP + lo + i], x2 = sm.Q[(3 * a + 2) * TP + lo + i];
      const T q0 = x0 * i00;
      const T q1 = (x1 - q0 * r01) * i11;
      const T q2 = (x2 - q0 * r02 - q1 * r12) * i22;
      sm.Q[(3 * a + 0) * TP + lo + i] = q0; sm.Q[(3 * a + 1) * TP + lo + i] = q1; sm.Q[(3 * a + 2) * TP + lo + i] = q2;
      const T e = sm.E[a * TP + lo + i];
      c0 += q0 * e; c1 += q1 * e; c2 += q2 * e;
    }
  }
  sm.C[p] = c0; sm.C[TP + p] = c1; sm.C[2 * TP + p] = c2;
}

// ---------------------------------------------------------------------------------------------
// Shared prologue of K2 and K45: phases 1-3. Returns through shared memory. `t` = lane in CTA.
// Jc of the lane's observation stays in registers (jc[18]).
// ---------------------------------------------------------------------------------------------
template <class T> struct TileArgs {
  const int* __restrict__ tile_pt;    // [ntiles+1] first point of each tile
  const int* __restrict__ pt_start;   // [M+1]
  const int* __restrict__ view;       // [K]
  const int* __restrict__ point;      // [K]
  const T* __restrict__ meas;         // [2K]
  const T* __restrict__ cams;         // [N*16]
  const T* __restrict__ X;            // [3M]
  T tau2;
  T lambda;
  int factor;                          // PointFactor
};

template <class T>
__device__ __forceinline__ void tile_phases_123(const TileArgs<T>& a, TileSmem<T>& sm, int t, int p0, int npts, int o0, int nobs,
                                                int& cam_idx, int& lp) {
  cam_idx = -1; lp = 0;
  if (t < npts) {
    const int s = a.pt_start[p0 + t];
    sm.ptObs0[t] = s - o0;
    sm.ptN[t] = a.pt_start[p0 + t + 1] - s;
  }
  if (t < nobs) {
    const int i = o0 + t;
    cam_idx = __ldg(a.view + i);
    const int pj = __ldg(a.point + i);
    lp = pj - p0;
    Cam<T> c; load_cam<T>(a.cams, cam_idx, c);
    const T X0 = __ldg(a.X + 3 * (size_t)pj), X1 = __ldg(a.X + 3 * (size_t)pj + 1), X2 = __ldg(a.X + 3 * (size_t)pj + 2);
    const T m0 = __ldg(a.meas + 2 * (size_t)i), m1 = __ldg(a.meas + 2 * (size_t)i + 1);
    T e0, e1, jc[18], jp[6];
    obs_jacobian<T>(c, X0, X1, X2, m0, m1, a.tau2, e0, e1, jc, jp);
    sm.E[t] = e0; sm.E[TP + t] = e1;
#pragma unroll
    for (int k = 0; k < 6; ++k) sm.Q[k * TP + t] = jp[k];
#pragma unroll
    for (int k = 0; k < 18; ++k) sm.Jc[k * TP + t] = jc[k];
  }
  __syncthreads();
  if (t < npts) {
    const int n = sm.ptN[t];
    if (a.factor == PF_HOUSEHOLDER && n >= 2) point_householder<T>(sm, t, sm.ptObs0[t], n, tsqrt(a.lambda));
    else point_normal<T>(sm, t, sm.ptObs0[t], n, a.lambda);
  }
  __syncthreads();
  if (t < nobs) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const T qa = sm.Q[k * TP + t], qb = sm.Q[(3 + k) * TP + t];
#pragma unroll
      for (int b = 0; b < 9; ++b) sm.R12[(9 * k + b) * TP + t] = qa * sm.Jc[b * TP + t] + qb * sm.Jc[(9 + b) * TP + t];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2: accumulate the reduced camera system. Sv/lds: band view of S (entry (i,j), j<=i, at
// Sv[i*lds + j]); g[9N].  S_ab += delta_ab Jc_a^T Jc_a - R12_a^T R12_b ; g_a += Jc_a^T e_a - R12_a^T c.
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(TILE) k_schur(TileArgs<T> a, T* __restrict__ Sv, size_t lds, T* __restrict__ g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<T>& sm = *reinterpret_cast<TileSmem<T>*>(smem_raw);
  const int t = threadIdx.x, tile = blockIdx.x;
  const int p0 = a.tile_pt[tile], p1 = a.tile_pt[tile + 1], npts = p1 - p0;
  const int o0 = a.pt_start[p0], nobs = a.pt_start[p1] - o0;
  int cam_idx, lp;
  tile_phases_123<T>(a, sm, t, p0, npts, o0, nobs, cam_idx, lp);
  // pair offsets
  if (t < npts) { const int n = sm.ptN[t]; sm.pairOff[t + 1] = n * (n + 1) / 2; }
  if (t == 0) sm.pairOff[0] = 0;
  __syncthreads();
  if (t == 0) { int s = 0; for (int p = 0; p < npts; ++p) { s += sm.pairOff[p + 1]; sm.pairOff[p + 1] = s; } }
  // g contributions (one lane per observation)
  if (t < nobs) {
    const T e0 = sm.E[t], e1 = sm.E[TP + t];
    const T c0 = sm.C[lp], c1 = sm.C[TP + lp], c2 = sm.C[2 * TP + lp];
#pragma unroll
    for (int b = 0; b < 9; ++b) {
      const T v = sm.Jc[b * TP + t] * e0 + sm.Jc[(9 + b) * TP + t] * e1
                - (sm.R12[b * TP + t] * c0 + sm.R12[(9 + b) * TP + t] * c1 + sm.R12[(18 + b) * TP + t] * c2);
      atomic_add<T>(g + 9 * (size_t)cam_idx + b, v);
    }
  }
  __syncthreads();
  // phase 4: one warp per (point, ia >= ib) pair; lanes <-> entries of the 9x9 block
  const int lane = t & 31, warp = t >> 5, nwarps = TILE / 32;
  const int npairs = sm.pairOff[npts];
  int ep[3], eq[3];
#pragma unroll
  for (int s = 0; s < 3; ++s) { const int e = lane + 32 * s; ep[s] = e / 9; eq[s] = e - 9 * ep[s]; }
  for (int q = warp; q < npairs; q += nwarps) {
    int lo_ = 0, hi_ = npts;  // largest p with pairOff[p] <= q
    while (hi_ - lo_ > 1) { const int mid = (lo_ + hi_) >> 1; if (sm.pairOff[mid] <= q) lo_ = mid; else hi_ = mid; }
    const int pp = lo_, ql = q - sm.pairOff[pp];
    int ia = (int)((sqrtf(8.0f * (float)ql + 1.0f) - 1.0f) * 0.5f);
    while ((ia + 1) * (ia + 2) / 2 <= ql) ++ia;
    while (ia * (ia + 1) / 2 > ql) --ia;
    const int ib = ql - ia * (ia + 1) / 2;
    const int oa = sm.ptObs0[pp] + ia, ob = sm.ptObs0[pp] + ib;
    const int ca = __ldg(a.view + o0 + oa), cb = __ldg(a.view + o0 + ob);  // ca >= cb (sorted by camera)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int e = lane + 32 * s;
      if (e >= 81) continue;
      const int p = ep[s], qq = eq[s];
      if (ia == ib && qq > p) continue;
      T v = -(sm.R12[p * TP + oa] * sm.R12[qq * TP + ob] + sm.R12[(9 + p) * TP + oa] * sm.R12[(9 + qq) * TP + ob] +
              sm.R12[(18 + p) * TP + oa] * sm.R12[(18 + qq) * TP + ob]);
      if (ia == ib) v += sm.Jc[p * TP + oa] * sm.Jc[qq * TP + oa] + sm.Jc[(9 + p) * TP + oa] * sm.Jc[(9 + qq) * TP + oa];
      atomic_add<T>(Sv + (size_t)(9 * ca + p) * lds + (9 * cb + qq), v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K45: back-substitution, state update and test-point energy, fused.
//   dx_j = R_j^-1 (-c_j - sum_i R12_i dx_cam(i)) un-permuted; X_test = X + dx_j;
//   e_test = residual(cams_test, X_test); partial sums per tile:
//     part[0] = sum e_test^2, part[1] = |dx_pts|^2, part[2] = sum_obs e.(J dx)  (so that
//     dx^T(lambda dx + JtRes) = lambda |dx|^2 - part[2]).
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(TILE) k_backsub_eval(TileArgs<T> a, const T* __restrict__ dx_cam, const T* __restrict__ cams_test,
                                                       T* __restrict__ dx_pt, T* __restrict__ X_test, double* __restrict__ partials, int ntiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<T>& sm = *reinterpret_cast<TileSmem<T>*>(smem_raw);
  const int t = threadIdx.x, tile = blockIdx.x;
  const int p0 = a.tile_pt[tile], p1 = a.tile_pt[tile + 1], npts = p1 - p0;
  const int o0 = a.pt_start[p0], nobs = a.pt_start[p1] - o0;
  int cam_idx, lp;
  tile_phases_123<T>(a, sm, t, p0, npts, o0, nobs, cam_idx, lp);
  double acc_e = 0.0, acc_dx = 0.0, acc_jd = 0.0;
  // per observation: u_i = R12_i dx_cam(i) (3) -> E rows reused after reading e; e.(Jc dx_cam)
  T u0 = T(0), u1 = T(0), u2 = T(0);
  if (t < nobs) {
    T d[9];
#pragma unroll
    for (int b = 0; b < 9; ++b) d[b] = __ldg(dx_cam + 9 * (size_t)cam_idx + b);
    T j0 = T(0), j1 = T(0);
#pragma unroll
    for (int b = 0; b < 9; ++b) {
      u0 += sm.R12[b * TP + t] * d[b]; u1 += sm.R12[(9 + b) * TP + t] * d[b]; u2 += sm.R12[(18 + b) * TP + t] * d[b];
      j0 += sm.Jc[b * TP + t] * d[b]; j1 += sm.Jc[(9 + b) * TP + t] * d[b];
    }
    acc_jd += (double)(sm.E[t] * j0 + sm.E[TP + t] * j1);
  }
  __syncthreads();  // all reads of Jc done; reuse Jc rows 0..2 as per-observation u
  if (t < nobs) { sm.Jc[0 * TP + t] = u0; sm.Jc[1 * TP + t] = u1; sm.Jc[2 * TP + t] = u2; }
  __syncthreads();
  if (t < npts) {
    const int lo = sm.ptObs0[t], n = sm.ptN[t];
    T r0 = -sm.C[t], r1 = -sm.C[TP + t], r2 = -sm.C[2 * TP + t];
    for (int i = 0; i < n; ++i) { r0 -= sm.Jc[0 * TP + lo + i]; r1 -= sm.Jc[1 * TP + lo + i]; r2 -= sm.Jc[2 * TP + lo + i]; }
    const T z2 = r2 / sm.Rm[5 * TP + t];
    const T z1 = (r1 - sm.Rm[4 * TP + t] * z2) / sm.Rm[3 * TP + t];
    const T z0 = (r0 - sm.Rm[1 * TP + t] * z1 - sm.Rm[2 * TP + t] * z2) / sm.Rm[0 * TP + t];
    const int pm = sm.perm[t];
    T d[3];
    d[pm & 3] = z0; d[(pm >> 2) & 3] = z1; d[(pm >> 4) & 3] = z2;
    const size_t gp = 3 * (size_t)(p0 + t);
    acc_dx += (double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    acc_jd += (double)(sm.G[t] * d[0] + sm.G[TP + t] * d[1] + sm.G[2 * TP + t] * d[2]);
    dx_pt[gp] = d[0]; dx_pt[gp + 1] = d[1]; dx_pt[gp + 2] = d[2];
    const T x0 = __ldg(a.X + gp) + d[0], x1 = __ldg(a.X + gp + 1) + d[1], x2 = __ldg(a.X + gp + 2) + d[2];
    X_test[gp] = x0; X_test[gp + 1] = x1; X_test[gp + 2] = x2;
    sm.G[t] = x0; sm.G[TP + t] = x1; sm.G[2 * TP + t] = x2;
  }
  __syncthreads();
  if (t < nobs) {
    Cam<T> c; load_cam<T>(cams_test, cam_idx, c);
    const int i = o0 + t;
    const T m0 = __ldg(a.meas + 2 * (size_t)i), m1 = __ldg(a.meas + 2 * (size_t)i + 1);
    T e0, e1;
    obs_residual<T>(c, sm.G[lp], sm.G[TP + lp], sm.G[2 * TP + lp], m0, m1, a.tau2, e0, e1);
    acc_e += (double)(e0 * e0 + e1 * e1);
  }
  // block reduction (deterministic order)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc_e += __shfl_down_sync(0xffffffffu, acc_e, off);
    acc_dx += __shfl_down_sync(0xffffffffu, acc_dx, off);
    acc_jd += __shfl_down_sync(0xffffffffu, acc_jd, off);
  }
  const int lane = t & 31, warp = t >> 5;
  if (lane == 0) { sm.red[warp] = acc_e; sm.red[TILE / 32 + warp] = acc_dx; sm.red[2 * (TILE / 32) + warp] = acc_jd; }
  __syncthreads();
  if (t < 3) {
    double s = 0.0;
    for (int w = 0; w < TILE / 32; ++w) s += sm.red[t * (TILE / 32) + w];
    partials[(size_t)t * ntiles + tile] = s;
  }
}

}  // namespace ba

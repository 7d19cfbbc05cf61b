// Point-tile kernels for sm_100a: the fused Jacobian + per-point block factorisation kernel
// (k_point_factor), the deterministic Schur gather (k_schur_gather) and the fused back-substitution +
// update + test-energy kernel (k_backsub_eval).
//
// Replaces, on the device and without ever materialising J:
//   row permutation + [J; sqrt(lambda) I]                  BacktrackLevMarqQRChol.h:291-315
//   BlockDiagonalSparseQR<.., ColPivHouseholderQR>::compute on the (2n_j+3)x3 point blocks and
//   Q1^T * J2 (camera columns), matrixQ().transpose()*r     QRChol.h:319-329 (solver NOT IN TREE)
//   J2bot^T J2bot, J2bot^T qtb2                             QRChol.h:339-341
//   back-substitution + colsPermutation                    QRChol.h:344-360
//   increment_in_place + functor(xTest)                    QRChol.h:363-371, BAFunctor.h:299-342
//
// Work decomposition (B200-first): k_point_factor_warp gives one WARP a unit of consecutive
// points whose observations fit 32 lanes (lane = observation, everything in registers, segmented shuffles);
// the tile kernel k_point_factor (a CTA owns consecutive points whose observations fit TILE lanes: phase 1 one
// lane per OBSERVATION, phase 2 one lane per POINT, phase 3 one lane per observation) only serves points with
// 33..128 observations, k_point_factor_big the longer tracks. All of them write per-observation records; the
// reduced camera matrix is then WRITTEN (no atomics) by k_schur_diag / k_schur_gather from static pair lists.
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
  double red[3 * (TILE / 32)];
  __device__ static constexpr int stride() { return TP; }
  __device__ static constexpr int ostride() { return TP; }
};
// One point per CTA (more than TILE observations): Q / E rows in dynamic shared memory with a runtime stride,
// the point's outputs (p = 0) in small arrays.
template <class T> struct BigPointStore {
  T* Q; T* E; int st;
  T Rm[6], C[3], G[3]; int perm[1];
  __device__ int stride() const { return st; }
  __device__ static constexpr int ostride() { return 1; }
};

// ---------------------------------------------------------------------------------------------
// Phase 2: one lane per point. Rows rho = 2*i + a live in sm.Q[(3a+b)*TP + lo + i]; the three
// lambda rows sqrt(lambda) I3 live in registers. Eigen ColPivHouseholderQR conventions.
// ---------------------------------------------------------------------------------------------
// The storage is a template parameter: TileSmem (stride TP, one lane per point) or BigPointStore (one point per CTA,
// runtime stride; points with more than TILE observations).
template <class T, class S>
__device__ __forceinline__ T& qel(S& sm, int lo, int rho, int b) { return sm.Q[(3 * (rho & 1) + b) * sm.stride() + lo + (rho >> 1)]; }

template <class T, class S>
__device__ void point_householder(S& sm, int p, int lo, int n, T sl) {
  const int m = 2 * n;  // observation rows; plus 3 register rows
  T L[3][3] = {{sl, T(0), T(0)}, {T(0), sl, T(0)}, {T(0), T(0), sl}};
  T tau[3] = {T(0), T(0), T(0)};
  T Rv[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  int pm0 = 0, pm1 = 1, pm2 = 2;
  // g_p = Jp^T e in the ORIGINAL column order
  {
    T g0 = T(0), g1 = T(0), g2 = T(0);
    for (int i = 0; i < n; ++i) {
      const T e0 = sm.E[lo + i], e1 = sm.E[sm.stride() + lo + i];
      g0 += sm.Q[0 * sm.stride() + lo + i] * e0 + sm.Q[3 * sm.stride() + lo + i] * e1;
      g1 += sm.Q[1 * sm.stride() + lo + i] * e0 + sm.Q[4 * sm.stride() + lo + i] * e1;
      g2 += sm.Q[2 * sm.stride() + lo + i] * e0 + sm.Q[5 * sm.stride() + lo + i] * e1;
    }
    sm.G[p] = g0; sm.G[sm.ostride() + p] = g1; sm.G[2 * sm.ostride() + p] = g2;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // squared norms of the remaining columns over rows >= k (pivot rule)
    T nn[3] = {T(-1), T(-1), T(-1)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < k) continue;
      T s = T(0);
      for (int r = k; r < m; ++r) { const T v = qel<T>(sm, lo, r, c); s += v * v; }
#pragma unroll
      for (int r = 0; r < 3; ++r) s += L[r][c] * L[r][c];
      nn[c] = s;
    }
    int best = k;
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k && nn[c] > nn[best]) best = c;
    if (best != k) {
      for (int r = 0; r < m; ++r) { T& a = qel<T>(sm, lo, r, k); T& b = qel<T>(sm, lo, r, best); const T t = a; a = b; b = t; }
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
    const T c0 = qel<T>(sm, lo, k, k);
    T tail2 = T(0);
    for (int r = k + 1; r < m; ++r) { const T v = qel<T>(sm, lo, r, k); tail2 += v * v; }
#pragma unroll
    for (int r = 0; r < 3; ++r) tail2 += L[r][k] * L[r][k];
    T beta, tk;
    if (tail2 <= (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f))) {
      tk = T(0); beta = c0;
      for (int r = k + 1; r < m; ++r) qel<T>(sm, lo, r, k) = T(0);
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][k] = T(0);
    } else {
      beta = tsqrt(c0 * c0 + tail2);
      if (c0 >= T(0)) beta = -beta;
      const T inv = T(1) / (c0 - beta);
      for (int r = k + 1; r < m; ++r) qel<T>(sm, lo, r, k) *= inv;
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][k] *= inv;
      tk = (beta - c0) / beta;
    }
    tau[k] = tk;
    qel<T>(sm, lo, k, k) = beta;
    // apply H_k to the remaining columns
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c <= k) continue;
      T s = qel<T>(sm, lo, k, c);
      for (int r = k + 1; r < m; ++r) s += qel<T>(sm, lo, r, k) * qel<T>(sm, lo, r, c);
#pragma unroll
      for (int r = 0; r < 3; ++r) s += L[r][k] * L[r][c];
      s *= tk;
      qel<T>(sm, lo, k, c) -= s;
      for (int r = k + 1; r < m; ++r) qel<T>(sm, lo, r, c) -= s * qel<T>(sm, lo, r, k);
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][c] -= s * L[r][k];
    }
  }
  // R (upper triangle) sits in observation rows 0..2 (n >= 2 => m >= 4); column swaps above were
  // applied to all rows, so it is already in pivoted column order.
  Rv[0] = qel<T>(sm, lo, 0, 0); Rv[1] = qel<T>(sm, lo, 0, 1); Rv[2] = qel<T>(sm, lo, 0, 2);
  Rv[3] = qel<T>(sm, lo, 1, 1); Rv[4] = qel<T>(sm, lo, 1, 2); Rv[5] = qel<T>(sm, lo, 2, 2);
#pragma unroll
  for (int i = 0; i < 6; ++i) sm.Rm[i * sm.ostride() + p] = Rv[i];
  sm.perm[p] = pm0 | (pm1 << 2) | (pm2 << 4);
  // form the thin Q1 in place (dorg2r): k = 2, 1, 0
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const T tk = tau[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c <= k) continue;
      T s = qel<T>(sm, lo, k, c);
      for (int r = k + 1; r < m; ++r) s += qel<T>(sm, lo, r, k) * qel<T>(sm, lo, r, c);
#pragma unroll
      for (int r = 0; r < 3; ++r) s += L[r][k] * L[r][c];
      s *= tk;
      qel<T>(sm, lo, k, c) -= s;
      for (int r = k + 1; r < m; ++r) qel<T>(sm, lo, r, c) -= s * qel<T>(sm, lo, r, k);
#pragma unroll
      for (int r = 0; r < 3; ++r) L[r][c] -= s * L[r][k];
    }
    for (int r = k + 1; r < m; ++r) qel<T>(sm, lo, r, k) *= -tk;
#pragma unroll
    for (int r = 0; r < 3; ++r) L[r][k] *= -tk;
    qel<T>(sm, lo, k, k) = T(1) - tk;
    for (int r = 0; r < k; ++r) qel<T>(sm, lo, r, k) = T(0);
  }
  // c = Q1^T e
  T c0 = T(0), c1 = T(0), c2 = T(0);
  for (int i = 0; i < n; ++i) {
    const T e0 = sm.E[lo + i], e1 = sm.E[sm.stride() + lo + i];
    c0 += sm.Q[0 * sm.stride() + lo + i] * e0 + sm.Q[3 * sm.stride() + lo + i] * e1;
    c1 += sm.Q[1 * sm.stride() + lo + i] * e0 + sm.Q[4 * sm.stride() + lo + i] * e1;
    c2 += sm.Q[2 * sm.stride() + lo + i] * e0 + sm.Q[5 * sm.stride() + lo + i] * e1;
  }
  sm.C[p] = c0; sm.C[sm.ostride() + p] = c1; sm.C[2 * sm.ostride() + p] = c2;
}

// Normal-equation point factor (CHOLESKY variant, BacktrackLevMarqCholesky.h:260-282):
// V = Jp^T Jp + lambda I = L D L^T; R := D^{1/2} L^T; Q1 rows := Jp rows * R^{-1}.
template <class T, class S>
__device__ void point_normal(S& sm, int p, int lo, int n, T lambda) {
  T v00 = lambda, v10 = T(0), v11 = lambda, v20 = T(0), v21 = T(0), v22 = lambda;
  T g0 = T(0), g1 = T(0), g2 = T(0);
  for (int i = 0; i < n; ++i) {
    const T a0 = sm.Q[0 * sm.stride() + lo + i], a1 = sm.Q[1 * sm.stride() + lo + i], a2 = sm.Q[2 * sm.stride() + lo + i];
    const T b0 = sm.Q[3 * sm.stride() + lo + i], b1 = sm.Q[4 * sm.stride() + lo + i], b2 = sm.Q[5 * sm.stride() + lo + i];
    const T e0 = sm.E[lo + i], e1 = sm.E[sm.stride() + lo + i];
    v00 += a0 * a0 + b0 * b0; v10 += a1 * a0 + b1 * b0; v11 += a1 * a1 + b1 * b1;
    v20 += a2 * a0 + b2 * b0; v21 += a2 * a1 + b2 * b1; v22 += a2 * a2 + b2 * b2;
    g0 += a0 * e0 + b0 * e1; g1 += a1 * e0 + b1 * e1; g2 += a2 * e0 + b2 * e1;
  }
  sm.G[p] = g0; sm.G[sm.ostride() + p] = g1; sm.G[2 * sm.ostride() + p] = g2;
  // V = Jp^T Jp + lambda I has no eigenvalue below lambda, hence no exact pivot below lambda either: the floor only
  // acts when cancellation (low-parallax points, float) has driven a computed pivot to or below zero, and keeps the
  // square roots below finite (the reference's SimplicialLDLT never takes the root of a pivot)
  const T d0 = v00, l10 = v10 / d0, l20 = v20 / d0;
  const T d1 = tmax(v11 - l10 * l10 * d0, lambda), l21 = (v21 - l20 * l10 * d0) / d1;
  const T d2 = tmax(v22 - l20 * l20 * d0 - l21 * l21 * d1, lambda);
  const T s0 = tsqrt(d0), s1 = tsqrt(d1), s2 = tsqrt(d2);
  const T r00 = s0, r01 = s0 * l10, r02 = s0 * l20, r11 = s1, r12 = s1 * l21, r22 = s2;
  sm.Rm[0 * sm.ostride() + p] = r00; sm.Rm[1 * sm.ostride() + p] = r01; sm.Rm[2 * sm.ostride() + p] = r02;
  sm.Rm[3 * sm.ostride() + p] = r11; sm.Rm[4 * sm.ostride() + p] = r12; sm.Rm[5 * sm.ostride() + p] = r22;
  sm.perm[p] = 0 | (1 << 2) | (2 << 4);
  const T i00 = T(1) / r00, i11 = T(1) / r11, i22 = T(1) / r22;
  T c0 = T(0), c1 = T(0), c2 = T(0);
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const T x0 = sm.Q[(3 * a + 0) * sm.stride() + lo + i], x1 = sm.Q[(3 * a + 1) * sm.stride() + lo + i], x2 = sm.Q[(3 * a + 2) * sm.stride() + lo + i];
      const T q0 = x0 * i00;
      const T q1 = (x1 - q0 * r01) * i11;
      const T q2 = (x2 - q0 * r02 - q1 * r12) * i22;
      sm.Q[(3 * a + 0) * sm.stride() + lo + i] = q0; sm.Q[(3 * a + 1) * sm.stride() + lo + i] = q1; sm.Q[(3 * a + 2) * sm.stride() + lo + i] = q2;
      const T e = sm.E[a * sm.stride() + lo + i];
      c0 += q0 * e; c1 += q1 * e; c2 += q2 * e;
    }
  }
  sm.C[p] = c0; sm.C[sm.ostride() + p] = c1; sm.C[2 * sm.ostride() + p] = c2;
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
    if (a.factor == PF_HOUSEHOLDER && n >= 2) point_householder<T, TileSmem<T>>(sm, t, sm.ptObs0[t], n, tsqrt(a.lambda));
    else point_normal<T, TileSmem<T>>(sm, t, sm.ptObs0[t], n, a.lambda);
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

// =============================================================================================
// Deterministic two-pass Schur accumulation (no atomics).
//
// FP64 global reductions issue at ~1.3 cycles per lane per SM (tools/ubench; 1.34e9 of them = 6 ms at the
// synthetic scale) and make S differ from run to run. The structure of S is static (which point
// contributes to which 9x9 camera-pair block), so it is built once on the host: every non-empty block
// (a, b), b <= a, owns the list of (observation of a, observation of b) pairs of the points seen by both.
// Pass 1 (k_point_factor_warp / k_point_factor) stores per observation two records of REC = 28 scalars
// (224 / 112 bytes):
//   P record, at the observation's index (point-major, so the back-substitution streams them):
//             R12_i = Q1_i^T Jc_i, 3x9 row-major, one pad scalar       (off-diagonal blocks, back-substitution)
//   D record, at the observation's CAMERA-MAJOR slot (observations sorted by (camera, point), so the diagonal
//             kernel streams a camera's records): Jc_i 2x9 row-major | Q1_i 2x3 (the observation's rows of the thin Q) |
//             w_i = e_i - Q1_i c | e_i
//             (diagonal blocks: Jc^T Jc - R12^T R12 = Jc^T M Jc with M_i = I2 - Q1_i Q1_i^T; g: Jc^T e - R12^T c = Jc^T w;
//             gJ = Jc^T e; Q1_i also serves the corrected semi-normal refinement of the QR variants, k_csne_*)
// and per point (R (6), c (3), G = Jp^T e (3), perm, pad) = 16 scalars.
// Pass 2: k_schur_gather (one warp per off-diagonal block, one LANE per pair with the whole 9x9 block in
// registers, records staged through shared memory with cp.async one batch of 32 pairs ahead) and
// k_schur_diag (one CTA per camera streaming the camera's contiguous D records, one lane per record)
// WRITE the blocks: every entry of S has exactly one writer and a fixed summation order, so the result is
// bit-reproducible. Blocks are processed in (a, b) order, so the P records of the ~bw cameras in flight stay in
// L2 whatever their storage order. The back-substitution re-reads the P and point records instead of
// re-evaluating the Jacobian and the point QR.
//
// Why one lane per pair: the LSU delivers 128 bytes per clock per SM to the register file, and every
// distinct 128-byte line touched by a warp instruction costs one pass. With 3x3 register tiles (9 lanes per
// pair, r1 v3) each lane re-loads 6 operands per 9 FMAs from three different records per instruction: ncu
// showed l1tex at 96 % and 25 passes per pair. A lane that owns the whole block needs 18 operands per 81
// FMAs, and the staged records are read conflict-free (stride 30 doubles / 36 floats, 16-byte loads).
// =============================================================================================
constexpr int REC = 28;   // scalars per observation record (P and D)
constexpr int PREC = 16;  // scalars per point record
// geometry of a record staged in shared memory (see "Staging of records" below)
template <class T> struct RecGeom {
  static constexpr int EPC = 16 / (int)sizeof(T);      // scalars per 16-byte chunk
  static constexpr int CPR = REC / EPC;                // chunks per record (14 / 7)
  static constexpr int SREC = sizeof(T) == 8 ? 30 : 36;
};

__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
  *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
}
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void load4(const double* p, double& a, double& b, double& c, double& d) {
  const double2 u = __ldg(reinterpret_cast<const double2*>(p)), v = __ldg(reinterpret_cast<const double2*>(p + 2));
  a = u.x; b = u.y; c = v.x; d = v.y;
}
__device__ __forceinline__ void load4(const float* p, float& a, float& b, float& c, float& d) {
  const float4 u = __ldg(reinterpret_cast<const float4*>(p));
  a = u.x; b = u.y; c = u.z; d = u.w;
}
template <class T> __device__ __forceinline__ void store_rec(T* p, const T (&v)[REC]) {
#pragma unroll
  for (int c = 0; c < REC; c += 4) store4(p + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
}
// D record tail from the observation's thin-Q rows q[a][k], c = Q1^T e of its point and its residual
template <class T>
__device__ __forceinline__ void fill_drec_tail(T (&rec)[REC], const T q00, const T q01, const T q02, const T q10, const T q11, const T q12,
                                               const T c0, const T c1, const T c2, const T e0, const T e1) {
  rec[18] = q00; rec[19] = q01; rec[20] = q02; rec[21] = q10; rec[22] = q11; rec[23] = q12;
  rec[24] = e0 - (q00 * c0 + q01 * c1 + q02 * c2);
  rec[25] = e1 - (q10 * c0 + q11 * c1 + q12 * c2);
  rec[26] = e0; rec[27] = e1;
}

template <class T>
__global__ void __launch_bounds__(TILE) k_point_factor(TileArgs<T> a, const int stride, const int* __restrict__ slot, T* __restrict__ Prec,
                                                       T* __restrict__ Drec, T* __restrict__ Ptrec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<T>& sm = *reinterpret_cast<TileSmem<T>*>(smem_raw);
  const int t = threadIdx.x, tile = blockIdx.x;
  const int p0 = a.tile_pt[stride * tile], p1 = a.tile_pt[stride * tile + 1], npts = p1 - p0;  // stride 2: (first, end) pairs
  const int o0 = a.pt_start[p0], nobs = a.pt_start[p1] - o0;
  int cam_idx, lp;
  tile_phases_123<T>(a, sm, t, p0, npts, o0, nobs, cam_idx, lp);
  __syncthreads();
  if (t < nobs) {
    const size_t sl = (size_t)__ldg(slot + o0 + t);
    const T e0 = sm.E[t], e1 = sm.E[TP + t];
    T rec[REC];
#pragma unroll
    for (int b = 0; b < 27; ++b) rec[b] = sm.R12[b * TP + t];
    rec[27] = T(0);
    store_rec(Prec + (size_t)(o0 + t) * REC, rec);
#pragma unroll
    for (int b = 0; b < 18; ++b) rec[b] = sm.Jc[b * TP + t];
    fill_drec_tail<T>(rec, sm.Q[0 * TP + t], sm.Q[1 * TP + t], sm.Q[2 * TP + t], sm.Q[3 * TP + t], sm.Q[4 * TP + t], sm.Q[5 * TP + t],
                      sm.C[lp], sm.C[TP + lp], sm.C[2 * TP + lp], e0, e1);
    store_rec(Drec + sl * REC, rec);
  }
  if (t < npts) {
    T* q = Ptrec + (size_t)(p0 + t) * PREC;
    store4(q, sm.Rm[0 * TP + t], sm.Rm[1 * TP + t], sm.Rm[2 * TP + t], sm.Rm[3 * TP + t]);
    store4(q + 4, sm.Rm[4 * TP + t], sm.Rm[5 * TP + t], sm.C[t], sm.C[TP + t]);
    store4(q + 8, sm.C[2 * TP + t], sm.G[t], sm.G[TP + t], sm.G[2 * TP + t]);
    store4(q + 12, (T)sm.perm[t], T(0), T(0), T(0));
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-segmented point factor: one WARP per unit of consecutive points whose observations fit 32 lanes,
// lane = observation, everything in registers. A point's rows live in the lanes of its segment; sums
// over a point (column norms, reflector dots, Q1^T e, Jp^T e, normal equations) are segmented sums done
// with shuffles in row order (every lane of the segment computes the identical value, so the 3x3 R,
// the reflector scalars and the sqrt(lambda) I3 rows are simply replicated in the segment's lanes).
// No shared memory, no block barriers: the tile version (k_point_factor, kept for points with more than
// 32 observations) spends most of its time in one-lane-per-point loops over shared memory behind
// __syncthreads (ncu: barrier stall 17.7 per issue, 9 % issue utilisation).
// ---------------------------------------------------------------------------------------------
// Segmented sum over the lanes of a point (segment = lanes s0 .. s0+n-1, i = lane - s0): Hillis-Steele scan with
// shfl_up restricted to the segment, then the last lane's total is broadcast. Every lane of a segment ends up with
// the identical value; fixed order -> deterministic. Only entries [lo, hi) are summed (compile-time pruned).
template <class T, int NV>
__device__ __forceinline__ void seg_sum(T (&v)[NV], const int s0, const int n, const int nmax, const int i, const int lo = 0, const int hi = NV) {
  constexpr unsigned FULL = 0xffffffffu;
  for (int d = 1; d < nmax; d <<= 1) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      if (q >= lo && q < hi) { const T o = __shfl_up_sync(FULL, v[q], d); if (i >= d) v[q] += o; }
    }
  }
  const int lastl = (s0 + n - 1) & 31;
#pragma unroll
  for (int q = 0; q < NV; ++q) if (q >= lo && q < hi) v[q] = __shfl_sync(FULL, v[q], lastl);
}
template <class T> __device__ __forceinline__ void cswap(T& a, T& b, bool sw) { const T t = a; a = sw ? b : a; b = sw ? t : b; }

// x[a][c]: the lane's two rows of Jp (row 2i+a of the point block); on return the thin Q1 rows.
template <class T>
__device__ __forceinline__ void seg_householder(T (&x)[2][3], const T e0, const T e1, const T sl, const int s0, const int n, const int i,
                                                const int nmax, T (&R)[6], int& pm, T (&cq)[3]) {
  constexpr unsigned FULL = 0xffffffffu;
  const T tiny = (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f));
  T L[3][3] = {{sl, T(0), T(0)}, {T(0), sl, T(0)}, {T(0), T(0), sl}};
  T tau[3];
  int pmv[3] = {0, 1, 2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const bool ge0 = 2 * i >= k, ge1 = 2 * i + 1 >= k, gt0 = 2 * i > k, gt1 = 2 * i + 1 > k;
    const int own = (s0 + (k >> 1)) & 31;  // lane holding row k, in slot k & 1
    // squared norms of the remaining columns over rows >= k (pivot rule)
    T nn[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c >= k) nn[c] = (ge0 ? x[0][c] * x[0][c] : T(0)) + (ge1 ? x[1][c] * x[1][c] : T(0));
    seg_sum<T, 3>(nn, s0, n, nmax, i, k);
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c >= k) nn[c] += L[0][c] * L[0][c] + L[1][c] * L[1][c] + L[2][c] * L[2][c];
    int best = k;
    T bestv = nn[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k && nn[c] > bestv) { best = c; bestv = nn[c]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) {
        const bool sw = best == c;
        cswap(x[0][k], x[0][c], sw); cswap(x[1][k], x[1][c], sw);
        cswap(L[0][k], L[0][c], sw); cswap(L[1][k], L[1][c], sw); cswap(L[2][k], L[2][c], sw);
        const int t = pmv[k]; pmv[k] = sw ? pmv[c] : pmv[k]; pmv[c] = sw ? t : pmv[c];
      }
    }
    const T c0 = __shfl_sync(FULL, (k & 1) ? x[1][k] : x[0][k], own);
    T t2[1] = {(gt0 ? x[0][k] * x[0][k] : T(0)) + (gt1 ? x[1][k] * x[1][k] : T(0))};
    seg_sum<T, 1>(t2, s0, n, nmax, i);
    const T tail2 = t2[0] + L[0][k] * L[0][k] + L[1][k] * L[1][k] + L[2][k] * L[2][k];
    const bool degenerate = tail2 <= tiny;
    T beta = tsqrt(c0 * c0 + tail2);
    if (c0 >= T(0)) beta = -beta;
    if (degenerate) beta = c0;
    const T inv = degenerate ? T(0) : T(1) / (c0 - beta);
    const T tk = degenerate ? T(0) : (beta - c0) / beta;
    if (gt0) x[0][k] *= inv;
    if (gt1) x[1][k] *= inv;
    L[0][k] *= inv; L[1][k] *= inv; L[2][k] *= inv;
    tau[k] = tk;
    const bool mine = (i == (k >> 1));
    if (mine) { if (k & 1) x[1][k] = beta; else x[0][k] = beta; }
    // apply H_k to the remaining columns
    T dots[2] = {T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k) dots[c - k - 1] = (gt0 ? x[0][k] * x[0][c] : T(0)) + (gt1 ? x[1][k] * x[1][c] : T(0));
    if (k < 2) seg_sum<T, 2>(dots, s0, n, nmax, i, 0, 2 - k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) {
        const T akc = __shfl_sync(FULL, (k & 1) ? x[1][c] : x[0][c], own);
        T sdot = akc + dots[c - k - 1] + L[0][k] * L[0][c] + L[1][k] * L[1][c] + L[2][k] * L[2][c];
        sdot *= tk;
        if (mine) { if (k & 1) x[1][c] -= sdot; else x[0][c] -= sdot; }
        if (gt0) x[0][c] -= sdot * x[0][k];
        if (gt1) x[1][c] -= sdot * x[1][k];
        L[0][c] -= sdot * L[0][k]; L[1][c] -= sdot * L[1][k]; L[2][c] -= sdot * L[2][k];
      }
    }
  }
  // R sits in rows 0..2 (lane s0: rows 0, 1; lane s0+1: row 2), already in pivoted column order
  const int l0 = s0 & 31, l1 = (s0 + 1) & 31;
  R[0] = __shfl_sync(FULL, x[0][0], l0); R[1] = __shfl_sync(FULL, x[0][1], l0); R[2] = __shfl_sync(FULL, x[0][2], l0);
  R[3] = __shfl_sync(FULL, x[1][1], l0); R[4] = __shfl_sync(FULL, x[1][2], l0); R[5] = __shfl_sync(FULL, x[0][2], l1);
  pm = pmv[0] | (pmv[1] << 2) | (pmv[2] << 4);
  // form the thin Q1 in place (dorg2r): k = 2, 1, 0
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const bool gt0 = 2 * i > k, gt1 = 2 * i + 1 > k, lt0 = 2 * i < k, lt1 = 2 * i + 1 < k;
    const int own = (s0 + (k >> 1)) & 31;
    const bool mine = (i == (k >> 1));
    const T tk = tau[k];
    T dots[2] = {T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k) dots[c - k - 1] = (gt0 ? x[0][k] * x[0][c] : T(0)) + (gt1 ? x[1][k] * x[1][c] : T(0));
    if (k < 2) seg_sum<T, 2>(dots, s0, n, nmax, i, 0, 2 - k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) {
        const T akc = __shfl_sync(FULL, (k & 1) ? x[1][c] : x[0][c], own);
        T sdot = akc + dots[c - k - 1] + L[0][k] * L[0][c] + L[1][k] * L[1][c] + L[2][k] * L[2][c];
        sdot *= tk;
        if (mine) { if (k & 1) x[1][c] -= sdot; else x[0][c] -= sdot; }
        if (gt0) x[0][c] -= sdot * x[0][k];
        if (gt1) x[1][c] -= sdot * x[1][k];
        L[0][c] -= sdot * L[0][k]; L[1][c] -= sdot * L[1][k]; L[2][c] -= sdot * L[2][k];
      }
    }
    if (gt0) x[0][k] *= -tk;
    if (gt1) x[1][k] *= -tk;
    L[0][k] *= -tk; L[1][k] *= -tk; L[2][k] *= -tk;
    if (mine) { if (k & 1) x[1][k] = T(1) - tk; else x[0][k] = T(1) - tk; }
    if (lt0) x[0][k] = T(0);
    if (lt1) x[1][k] = T(0);
  }
  cq[0] = x[0][0] * e0 + x[1][0] * e1; cq[1] = x[0][1] * e0 + x[1][1] * e1; cq[2] = x[0][2] * e0 + x[1][2] * e1;
  seg_sum<T, 3>(cq, s0, n, nmax, i);
}

// Normal-equation point factor (CHOLESKY variant; single-observation points): V = Jp^T Jp + lambda I = L D L^T,
// R := D^{1/2} L^T, Q1 rows := Jp rows * R^{-1}.
template <class T>
__device__ __forceinline__ void seg_normal(T (&x)[2][3], const T e0, const T e1, const T lambda, const int s0, const int n, const int i, const int nmax,
                                           T (&R)[6], int& pm, T (&cq)[3]) {
  T v[6] = {x[0][0] * x[0][0] + x[1][0] * x[1][0], x[0][1] * x[0][0] + x[1][1] * x[1][0], x[0][1] * x[0][1] + x[1][1] * x[1][1],
            x[0][2] * x[0][0] + x[1][2] * x[1][0], x[0][2] * x[0][1] + x[1][2] * x[1][1], x[0][2] * x[0][2] + x[1][2] * x[1][2]};
  seg_sum<T, 6>(v, s0, n, nmax, i);
  const T v00 = v[0] + lambda, v10 = v[1], v11 = v[2] + lambda, v20 = v[3], v21 = v[4], v22 = v[5] + lambda;
  const T d0 = v00, l10 = v10 / d0, l20 = v20 / d0;
  const T d1 = tmax(v11 - l10 * l10 * d0, lambda), l21 = (v21 - l20 * l10 * d0) / d1;  // pivot floor: see point_normal
  const T d2 = tmax(v22 - l20 * l20 * d0 - l21 * l21 * d1, lambda);
  const T q0 = tsqrt(d0), q1 = tsqrt(d1), q2 = tsqrt(d2);
  R[0] = q0; R[1] = q0 * l10; R[2] = q0 * l20; R[3] = q1; R[4] = q1 * l21; R[5] = q2;
  pm = 0 | (1 << 2) | (2 << 4);
  const T i00 = T(1) / R[0], i11 = T(1) / R[3], i22 = T(1) / R[5];
#pragma unroll
  for (int a2 = 0; a2 < 2; ++a2) {
    const T y0 = x[a2][0] * i00;
    const T y1 = (x[a2][1] - y0 * R[1]) * i11;
    const T y2 = (x[a2][2] - y0 * R[2] - y1 * R[4]) * i22;
    x[a2][0] = y0; x[a2][1] = y1; x[a2][2] = y2;
  }
  cq[0] = x[0][0] * e0 + x[1][0] * e1; cq[1] = x[0][1] * e0 + x[1][1] * e1; cq[2] = x[0][2] * e0 + x[1][2] * e1;
  seg_sum<T, 3>(cq, s0, n, nmax, i);
}

template <class T> constexpr size_t point_factor_warp_smem_bytes() { return (size_t)TILE * RecGeom<T>::SREC * sizeof(T); }
#ifndef BA_PF_MIN_BLOCKS
#define BA_PF_MIN_BLOCKS 4
#endif
// STAGE1 (MOREQR, BacktrackLevMarqMore.h:288-291: QR of the UN-damped J once per outer iteration): the kernel is run
// with lambda = 0 and stores, instead of the P / D records of a trial, one J record per observation at the observation's
// index (passed as `Prec`): Jc_i 2x9 | thin Q0 rows of the observation 2x3 | e_i | 2 pad, and the point record
// (R0, c0 = Q0^T e, G, perm0) into `Ptrec`; k_moreqr_stage2 turns them into the records of a trial.
template <class T, bool STAGE1 = false>
__global__ void __launch_bounds__(TILE, BA_PF_MIN_BLOCKS) k_point_factor_warp(TileArgs<T> a, int nunits, const int* __restrict__ unit_obs, const int* __restrict__ seg, const int* __restrict__ slot,
                                                            T* __restrict__ Prec, T* __restrict__ Drec, T* __restrict__ Ptrec) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, u = blockIdx.x * (TILE / 32) + (threadIdx.x >> 5);
  if (u >= nunits) return;
  // three levels of dependent loads (unit -> per-observation indices -> camera / point), not five: the unit table
  // holds (first observation, count) and `seg` holds each observation's (index within its point | count << 8)
  const int2 uo = __ldg(reinterpret_cast<const int2*>(unit_obs) + u);
  const int o0 = uo.x, un = uo.y;
  const bool act = lane < un;
  const int o = o0 + (act ? lane : 0);
  const int cam_idx = __ldg(a.view + o), pj = __ldg(a.point + o), sg = __ldg(seg + o);
  const size_t sl = (size_t)__ldg(slot + o);   // consumed at the very end: issued with the first level of loads
  const int i = act ? (sg & 0xff) : 0, n = act ? (sg >> 8) : 1;
  const int s0 = lane - i;
  Cam<T> c; load_cam<T>(a.cams, cam_idx, c);
  const T X0 = __ldg(a.X + 3 * (size_t)pj), X1 = __ldg(a.X + 3 * (size_t)pj + 1), X2 = __ldg(a.X + 3 * (size_t)pj + 2);
  const T m0 = __ldg(a.meas + 2 * (size_t)o), m1 = __ldg(a.meas + 2 * (size_t)o + 1);
  T e0, e1, jc[18], jp[6];
  obs_jacobian<T>(c, X0, X1, X2, m0, m1, a.tau2, e0, e1, jc, jp);
  T x[2][3] = {{jp[0], jp[1], jp[2]}, {jp[3], jp[4], jp[5]}};
  if (!act) {
    e0 = e1 = T(0);
#pragma unroll
    for (int q = 0; q < 3; ++q) { x[0][q] = T(0); x[1][q] = T(0); }
  }
  const int nmax = __reduce_max_sync(FULL, n);
  T G[3] = {x[0][0] * e0 + x[1][0] * e1, x[0][1] * e0 + x[1][1] * e1, x[0][2] * e0 + x[1][2] * e1};
  seg_sum<T, 3>(G, s0, n, nmax, i);
  const bool use_h = (a.factor == PF_HOUSEHOLDER) && n >= 2;
  T R[6], cq[3]; int pm = 0;
  T xh[2][3], Rh[6], ch[3]; int pmh = 0;
#pragma unroll
  for (int q = 0; q < 3; ++q) { xh[0][q] = x[0][q]; xh[1][q] = x[1][q]; }
  if (__any_sync(FULL, use_h)) seg_householder<T>(xh, e0, e1, tsqrt(a.lambda), s0, n, i, nmax, Rh, pmh, ch);
  if (__any_sync(FULL, !use_h)) seg_normal<T>(x, e0, e1, a.lambda, s0, n, i, nmax, R, pm, cq);
  if (use_h) {
#pragma unroll
    for (int q = 0; q < 3; ++q) { x[0][q] = xh[0][q]; x[1][q] = xh[1][q]; cq[q] = ch[q]; }
#pragma unroll
    for (int q = 0; q < 6; ++q) R[q] = Rh[q];
    pm = pmh;
  }
  // Records leave through the warp's shared-memory staging area so that the global stores are coalesced: a lane
  // writing its own 224-byte record touches 32 different lines per store instruction (896 LSU passes per record
  // type and warp; with both record types that was as long as the whole kernel), two whole records per
  // instruction (lane = (record, 16-byte chunk)) need 4-5.
  extern __shared__ __align__(16) unsigned char pfw_smem_raw[];
  constexpr int EPC = RecGeom<T>::EPC, CPR = RecGeom<T>::CPR, SR = RecGeom<T>::SREC;
  constexpr int RPI = 32 / CPR, NQ = 32 / RPI;
  T* const my = reinterpret_cast<T*>(pfw_smem_raw) + (size_t)(threadIdx.x >> 5) * 32 * SR;
  const int rsub = lane / CPR, part = lane - rsub * CPR;
  const bool cpl = lane < RPI * CPR;
  T rec[REC];
  if (STAGE1) {
#pragma unroll
    for (int b = 0; b < 18; ++b) rec[b] = jc[b];
#pragma unroll
    for (int k = 0; k < 3; ++k) { rec[18 + k] = x[0][k]; rec[21 + k] = x[1][k]; }
    rec[24] = e0; rec[25] = e1; rec[26] = T(0); rec[27] = T(0);
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int b = 0; b < 9; ++b) rec[9 * k + b] = x[0][k] * jc[b] + x[1][k] * jc[9 + b];
    rec[27] = T(0);
  }
  store_rec(my + lane * SR, rec);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int r = RPI * q + rsub;
    if (cpl && r < un)
      *reinterpret_cast<int4*>(Prec + (size_t)(o0 + r) * REC + part * EPC) = *reinterpret_cast<const int4*>(my + r * SR + part * EPC);
  }
  __syncwarp();
  if (!STAGE1) {
#pragma unroll
    for (int b = 0; b < 18; ++b) rec[b] = jc[b];
    fill_drec_tail<T>(rec, x[0][0], x[0][1], x[0][2], x[1][0], x[1][1], x[1][2], cq[0], cq[1], cq[2], e0, e1);
    store_rec(my + lane * SR, rec);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int r = RPI * q + rsub;
      const size_t slr = (size_t)__shfl_sync(FULL, (int)sl, r & 31);
      if (cpl && r < un)
        *reinterpret_cast<int4*>(Drec + slr * REC + part * EPC) = *reinterpret_cast<const int4*>(my + r * SR + part * EPC);
    }
  }
  if (!act) return;
  if (i == 0) {
    T* q = Ptrec + (size_t)pj * PREC;
    store4(q, R[0], R[1], R[2], R[3]);
    store4(q + 4, R[4], R[5], cq[0], cq[1]);
    store4(q + 8, cq[2], G[0], G[1], G[2]);
    store4(q + 12, (T)pm, T(0), T(0), T(0));
  }
}

// ---------------------------------------------------------------------------------------------
// Staging of records in shared memory. A record (REC scalars, 16-byte chunks) is copied with cp.async to a
// shared-memory slot of SREC scalars: 8 consecutive lanes reading 16 bytes each at that stride hit 32
// different banks (30 doubles = 60 words = 28 mod 32; 36 floats = 4 mod 32).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rec_cp16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
template <class T> __device__ __forceinline__ void rec_cp1(T* smem_dst, const T* gsrc) {  // one scalar
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void rec_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void rec_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Bulk asynchronous copy (the TMA engine's 1D mode: cp.async.bulk, UBLKCP in SASS) of a CONTIGUOUS run of records into
// shared memory, completion signalled on an mbarrier: one instruction by one lane replaces 14 (double) 16-byte cp.async per
// lane for a warp's 32 records. Source, destination and size must be multiples of 16 bytes (records are 224 / 112 bytes).
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(b), "r"(parity) : "memory");
}
// scalars [lo, hi) of a staged record into registers, 16-byte shared loads (lo, hi multiples of EPC)
template <int LO, int HI> __device__ __forceinline__ void lds_rec(const double* s, double (&r)[REC]) {
#pragma unroll
  for (int c = LO; c < HI; c += 2) { const double2 v = *reinterpret_cast<const double2*>(s + c); r[c] = v.x; r[c + 1] = v.y; }
}
template <int LO, int HI> __device__ __forceinline__ void lds_rec(const float* s, float (&r)[REC]) {
#pragma unroll
  for (int c = LO; c < HI; c += 4) { const float4 v = *reinterpret_cast<const float4*>(s + c); r[c] = v.x; r[c + 1] = v.y; r[c + 2] = v.z; r[c + 3] = v.w; }
}

// ---------------------------------------------------------------------------------------------
// Off-diagonal blocks S_ab = -sum_j R12_a^T R12_b: one warp per block (a, b), b < a, blocks handed out in
// (a, b) order through a global counter (the result does not depend on which warp computes a block). The
// block's pair list is cut into batches of 32 pairs; lane l owns pair l of the batch and accumulates its
// 9x9 product in 81 registers. While batch i is being multiplied, the 64 records of batch i+1 (possibly the
// first batch of the warp's next block) are in flight into the other half of the warp's staging buffer.
// At the end of a block the 32 lane accumulators are summed in lane order through shared memory (fixed
// order -> bit-reproducible) and lanes 0..26 write the block.
// ---------------------------------------------------------------------------------------------
#ifndef BA_GATHER_WARPS
#define BA_GATHER_WARPS 7
#endif
constexpr int GATHER_WARPS = BA_GATHER_WARPS;
constexpr int GATHER_THREADS = 32 * GATHER_WARPS;
template <class T> constexpr size_t gather_rec_bytes() { return (size_t)GATHER_WARPS * 2 * 64 * RecGeom<T>::SREC * sizeof(T); }
template <class T> constexpr size_t gather_smem_bytes() { return gather_rec_bytes<T>() + (size_t)GATHER_WARPS * 2 * 32 * sizeof(int2); }

template <class T>
__global__ void __launch_bounds__(GATHER_THREADS, 1) k_schur_gather(int nblocks, const int* __restrict__ blk_a, const int* __restrict__ blk_b,
                                                                    const int* __restrict__ blk_start, const int2* __restrict__ pairs,
                                                                    const T* __restrict__ Prec, T* __restrict__ Sv, size_t lds, int* __restrict__ counter) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int EPC = RecGeom<T>::EPC, CPR = RecGeom<T>::CPR, SR = RecGeom<T>::SREC;
  extern __shared__ __align__(16) unsigned char gather_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* const wbuf = reinterpret_cast<T*>(gather_smem_raw) + (size_t)warp * 2 * 64 * SR;
  // Three-deep software pipeline per warp. A warp issues in order, so a load costs its latency when its first
  // consumer is reached, and a register MOVE of a loaded value is a consumer (r1 v6: 30 % of all stall samples sat
  // on the first use of the pair indices, then on the loop-carried copy of them). Therefore nothing loaded is
  // carried in registers across iterations: while batch i is multiplied, the records of batch i+1 are in flight
  // into the other staging buffer and the pair indices of batch i+2 into a two-slot shared-memory ring (both
  // cp.async, one commit group per iteration); block ids are drawn from the global counter two blocks ahead and
  // the pair ranges one block ahead.
  struct Batch { int blk, start, cnt, ca, cb; bool last, valid; };
  int2* const ring = reinterpret_cast<int2*>(gather_smem_raw + (size_t)GATHER_WARPS * 2 * 64 * SR * sizeof(T)) + warp * 64;  // [2][32]
  int t = 0, tend = 0, blk = -1, ca = 0, cb = 0;   // current block: next pair, end of its pair range, id, cameras
  int id1 = 0, t1 = 0, end1 = 0, ca1 = 0, cb1 = 0; // next block (id resolved; range/cameras loading)
  int id2 = 0;                                     // block after next: counter value in flight (lane 0)
  auto draw = [&]() { if (lane == 0) id2 = atomicAdd(counter, 1); };
  auto load_next_block = [&]() {                   // consumes id2 (drawn one block ago), starts the range loads
    id1 = __shfl_sync(FULL, id2, 0);
    const int bc = min(id1, nblocks - 1);
    t1 = __ldg(blk_start + bc); end1 = __ldg(blk_start + bc + 1); ca1 = __ldg(blk_a + bc); cb1 = __ldg(blk_b + bc);
    draw();
  };
  auto next = [&]() -> Batch {
    Batch o; o.valid = false; o.blk = 0; o.start = 0; o.cnt = 0; o.last = false; o.ca = 0; o.cb = 0;
    if (t >= tend) {
      if (id1 >= nblocks) return o;
      blk = id1; t = t1; tend = end1; ca = ca1; cb = cb1;
      load_next_block();
    }
    o.valid = true; o.blk = blk; o.start = t; o.cnt = min(32, tend - t); t += o.cnt; o.last = (t >= tend); o.ca = ca; o.cb = cb;
    return o;
  };
  auto issue_pairs = [&](const Batch& bt, int2* slot) {
    if (bt.valid && lane < bt.cnt) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(slot + lane);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(pairs + bt.start + lane) : "memory");
    }
  };
  // records of a batch -> buf: slots 0..31 the a-records of the pairs, slots 32..63 the b-records. One
  // instruction copies RPI whole records (lane = (record, 16-byte chunk)), so the address arithmetic per copy
  // is one shuffle and one multiply-add.
  constexpr int RPI = 32 / CPR, NQ = 32 / RPI;
  const int rsub = lane / CPR, part = lane - rsub * CPR;
  const bool cpl = lane < RPI * CPR;
  const T* const src_lane = Prec + part * EPC;
  const int dst_lane = rsub * SR + part * EPC;
  auto issue_records = [&](const Batch& bt, const int2* slot, T* buf) {
    if (!bt.valid) return;
    int2 pr = make_int2(0, 0);
    if (lane < bt.cnt) pr = slot[lane];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int rec = RPI * q + rsub;
        const int sl = __shfl_sync(FULL, h ? pr.y : pr.x, rec & 31);
        if (cpl && rec < bt.cnt) rec_cp16(buf + dst_lane + (32 * h + RPI * q) * SR, src_lane + (size_t)sl * REC);
      }
    }
  };
  T acc[81];
#pragma unroll
  for (int e = 0; e < 81; ++e) acc[e] = T(0);
  draw();
  load_next_block();
  Batch c = next();
  issue_pairs(c, ring);
  rec_commit();
  Batch nx = next();
  rec_wait<0>();
  __syncwarp();
  issue_records(c, ring, wbuf);
  issue_pairs(nx, ring + 32);
  rec_commit();
  int cur = 0;
  while (c.valid) {
    rec_wait<0>();      // records of this batch and pair indices of the next one have landed
    __syncwarp();
    issue_records(nx, ring + 32 * (cur ^ 1), wbuf + (size_t)(cur ^ 1) * 64 * SR);
    const Batch nn = next();
    issue_pairs(nn, ring + 32 * cur);
    rec_commit();
    const int c_cnt = c.cnt; const bool c_last = c.last;
    T* const buf = wbuf + (size_t)cur * 64 * SR;
    if (lane < c_cnt) {
      const T* pa = buf + (size_t)lane * SR;
      const T* pb = buf + (size_t)(32 + lane) * SR;
      T ra[REC], rb[REC];
      // row k of R12 sits at scalars 9k .. 9k+8: stream the 16-byte chunks in three groups so that only the
      // operands of one row (plus the chunk straddling into the next) are live next to the 81 accumulators
      constexpr int G1 = (9 + EPC - 1) / EPC * EPC, G2 = (18 + EPC - 1) / EPC * EPC;
      lds_rec<0, G1>(pa, ra); lds_rec<0, G1>(pb, rb);
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[9 * i + j] += ra[i] * rb[j];
      lds_rec<G1, G2>(pa, ra); lds_rec<G1, G2>(pb, rb);
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[9 * i + j] += ra[9 + i] * rb[9 + j];
      lds_rec<G2, REC>(pa, ra); lds_rec<G2, REC>(pb, rb);
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[9 * i + j] += ra[18 + i] * rb[18 + j];
    }
    __syncwarp();
    if (c_last) {
      // fixed-order sum over the 32 lanes, 27 entries at a time, through the (consumed) staging buffer
      const int ca = c.ca, cb = c.cb;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
        for (int v = 0; v < 27; ++v) buf[v * 33 + lane] = acc[27 * ch + v];
        __syncwarp();
        if (lane < 27) {
          T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
          for (int l = 0; l < 32; l += 4) {
            s0 += buf[lane * 33 + l]; s1 += buf[lane * 33 + l + 1]; s2 += buf[lane * 33 + l + 2]; s3 += buf[lane * 33 + l + 3];
          }
          const int e = 27 * ch + lane, i = e / 9, j = e - 9 * i;
          Sv[(size_t)(9 * ca + i) * lds + 9 * cb + j] = -((s0 + s1) + (s2 + s3));
        }
        __syncwarp();
      }
#pragma unroll
      for (int e = 0; e < 81; ++e) acc[e] = T(0);
    }
    c = nx; nx = nn;
    cur ^= 1;
  }
  rec_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Diagonal blocks: one CTA per camera streams the camera's contiguous (camera-major) D records through
// shared memory (three cp.async stages of 32 records per warp) and every lane accumulates
//   S_aa += Jc^T M Jc (lower triangle, 45 entries),  g_a += Jc^T w,  gJ_a += Jc^T e
// for its record. Fixed-order reduction over lanes, then over warps (bit-reproducible).
// ---------------------------------------------------------------------------------------------
constexpr int DIAG_WARPS = 8;
constexpr int DIAG_THREADS = 32 * DIAG_WARPS;
constexpr int DIAG_STAGES = 3;
template <class T> constexpr size_t diag_smem_bytes() { return (size_t)DIAG_WARPS * DIAG_STAGES * 32 * RecGeom<T>::SREC * sizeof(T); }

template <class T>
__global__ void __launch_bounds__(DIAG_THREADS, 1) k_schur_diag(const int* __restrict__ cam_start, const T* __restrict__ Drec, T* __restrict__ Sv,
                                                                size_t lds, T* __restrict__ g, T* __restrict__ gJ, T lambda_diag) {
  constexpr int EPC = RecGeom<T>::EPC, CPR = RecGeom<T>::CPR, SR = RecGeom<T>::SREC;
  extern __shared__ __align__(16) unsigned char diag_smem_raw[];
  __shared__ T part[DIAG_WARPS][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ca = blockIdx.x;
  T* const wbuf = reinterpret_cast<T*>(diag_smem_raw) + (size_t)warp * DIAG_STAGES * 32 * SR;
  const int q0 = __ldg(cam_start + ca), q1 = __ldg(cam_start + ca + 1);
  const int nbatch = (q1 - q0 + 31) / 32;
  auto issue = [&](const int bi, T* buf) {
    if (bi < nbatch) {
      const int base = q0 + 32 * bi, cnt = min(32, q1 - base);
      const T* src = Drec + (size_t)base * REC;
#pragma unroll
      for (int q = 0; q < CPR; ++q) {
        const int idx = lane + 32 * q, rec = idx / CPR, part_ = idx - rec * CPR;
        if (rec < cnt) rec_cp16(buf + (size_t)rec * SR + part_ * EPC, src + (size_t)idx * EPC);
      }
    }
    rec_commit();
  };
  T acc[45], ga[9], gj[9];
#pragma unroll
  for (int e = 0; e < 45; ++e) acc[e] = T(0);
#pragma unroll
  for (int b = 0; b < 9; ++b) { ga[b] = T(0); gj[b] = T(0); }
  issue(warp, wbuf);
  issue(warp + DIAG_WARPS, wbuf + 32 * SR);
  int st = 0;
  for (int bi = warp; bi < nbatch; bi += DIAG_WARPS) {
    int st2 = st + 2; if (st2 >= DIAG_STAGES) st2 -= DIAG_STAGES;
    issue(bi + 2 * DIAG_WARPS, wbuf + (size_t)st2 * 32 * SR);
    rec_wait<2>();
    __syncwarp();
    const int cnt = min(32, q1 - (q0 + 32 * bi));
    if (lane < cnt) {
      T r[REC];
      lds_rec<0, REC>(wbuf + ((size_t)st * 32 + lane) * SR, r);
      // M_i = I2 - Q1_i Q1_i^T from the observation's thin-Q rows
      const T m00 = T(1) - (r[18] * r[18] + r[19] * r[19] + r[20] * r[20]);
      const T m01 = -(r[18] * r[21] + r[19] * r[22] + r[20] * r[23]);
      const T m11 = T(1) - (r[21] * r[21] + r[22] * r[22] + r[23] * r[23]);
      const T w0 = r[24], w1 = r[25], e0 = r[26], e1 = r[27];
      T t0[9], t1[9];
#pragma unroll
      for (int b = 0; b < 9; ++b) {
        t0[b] = m00 * r[b] + m01 * r[9 + b];
        t1[b] = m01 * r[b] + m11 * r[9 + b];
        ga[b] += r[b] * w0 + r[9 + b] * w1;
        gj[b] += r[b] * e0 + r[9 + b] * e1;
      }
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j)
          if (j <= i) acc[i * (i + 1) / 2 + j] += r[i] * t0[j] + r[9 + i] * t1[j];
    }
    __syncwarp();
    st = st + 1; if (st >= DIAG_STAGES) st = 0;
  }
  rec_wait<0>();
  __syncwarp();
  // lanes -> warp sums (21 values at a time through the warp's staging buffer), then warps in fixed order
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
    for (int v = 0; v < 21; ++v) {
      const int e = 21 * ch + v;
      wbuf[v * 33 + lane] = (e < 45) ? acc[e < 45 ? e : 0] : (e < 54 ? ga[(e >= 45 && e < 54) ? e - 45 : 0] : gj[e >= 54 ? e - 54 : 0]);
    }
    __syncwarp();
    if (lane < 21) {
      T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
      for (int l = 0; l < 32; l += 4) {
        s0 += wbuf[lane * 33 + l]; s1 += wbuf[lane * 33 + l + 1]; s2 += wbuf[lane * 33 + l + 2]; s3 += wbuf[lane * 33 + l + 3];
      }
      part[warp][21 * ch + lane] = (s0 + s1) + (s2 + s3);
    }
    __syncwarp();
  }
  __syncthreads();
  const int e = threadIdx.x;
  if (e < 63) {
    T v = T(0);
#pragma unroll
    for (int w = 0; w < DIAG_WARPS; ++w) v += part[w][e];
    if (e < 45) {
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      const int j = e - i * (i + 1) / 2;
      Sv[(size_t)(9 * ca + i) * lds + 9 * ca + j] = (i == j) ? v + lambda_diag : v;
    } else if (e < 54) {
      g[9 * ca + e - 45] = v;
    } else {
      gJ[9 * ca + e - 54] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Back-substitution, state update and test-point energy, fused, from the stored records:
//   dx_j = R_j^-1 (-c_j - sum_i R12_i dx_cam(i)) un-permuted; X_test = X + dx_j;
//   e_test = residual(cams_test, X_test); partial sums per tile:
//     part[0] = sum e_test^2, part[1] = |dx_pts|^2, part[2] = sum_points G . dx_j (G = Jp^T e), so that
//     dx^T(lambda dx + JtRes) = lambda |dx|^2 - part[2] - dx_cam . gJ.
// The tile's P records (scattered camera-major slots) and the dx_cam rows of its observations are staged
// with coalesced cp.async (consecutive lanes copy consecutive chunks of a record); a lane reading its own
// 224-byte record and 9 dx_cam scalars straight from global memory costs one LSU pass per lane per
// instruction (r1 v4: 27 passes per observation, l1tex 93 %).
// ---------------------------------------------------------------------------------------------
template <class T> constexpr size_t backsub_smem_bytes() { return (size_t)TILE * (RecGeom<T>::SREC + 9) * sizeof(T); }

template <class T>
__global__ void __launch_bounds__(TILE, 4) k_backsub_eval(TileArgs<T> a, const T* __restrict__ Prec,
                                                          const T* __restrict__ Ptrec, const T* __restrict__ dx_cam, const T* __restrict__ cams_test,
                                                          T* __restrict__ dx_pt, T* __restrict__ X_test, double* __restrict__ partials, int ntiles) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int EPC = RecGeom<T>::EPC, CPR = RecGeom<T>::CPR, SR = RecGeom<T>::SREC;
  constexpr int RPI = 32 / CPR, NQ = 32 / RPI;
  extern __shared__ __align__(16) unsigned char backsub_smem_raw[];
  T* const sP = reinterpret_cast<T*>(backsub_smem_raw);  // [TILE][SR]
  T* const sD = sP + (size_t)TILE * SR;                  // [TILE][9]
  __shared__ T su[3][TP];
  __shared__ T sx[3][TP];
  __shared__ double red[3 * (TILE / 32)];
  const int t = threadIdx.x, tile = blockIdx.x, lane = t & 31, w0 = t & ~31;
  const int4 ti = __ldg(reinterpret_cast<const int4*>(a.tile_pt) + tile);  // (first point, end point, first observation, observations)
  const int p0 = ti.x, p1 = ti.y, npts = p1 - p0, o0 = ti.z, nobs = ti.w;
  // Every warp stages the records and dx_cam rows of ITS 32 observations (lane = (record, chunk): one copy
  // instruction moves RPI whole records; the tile's records are contiguous), so a __syncwarp suffices before
  // they are read. Every load that does not depend on staged data is issued before the wait: the kernel is a
  // chain of dependent global loads (tile -> point range -> cameras -> dx_cam) and nothing else may add a level.
#ifdef BA_BULK_COPY
  // Experiment kept for the record: the warp's 32 records are contiguous in global memory, so ONE bulk
  // copy (TMA engine's 1D mode, cp.async.bulk -> UBLKCP + mbarrier) per warp can replace 14 16-byte cp.async per lane.
  // Measured on B200 at the synthetic scale: 0.457 ms (dense copy; the 16-byte record reads at stride 224 bytes then have
  // two-way bank conflicts) and 0.494 ms (one 224-byte bulk copy per record into the conflict-free stride) against
  // 0.433 ms for the LDGSTS path below, which therefore stays the default.
  constexpr int PS = REC;
  __shared__ unsigned long long bars[TILE / 32];
  const int wrec = min(32, nobs - w0);
  if (wrec > 0) {
    if (lane == 0) {
      mbar_init(&bars[t >> 5], 1);
      bulk_load(sP + (size_t)w0 * PS, Prec + (size_t)(o0 + w0) * REC, (unsigned)(wrec * REC * sizeof(T)), &bars[t >> 5]);
    }
    __syncwarp();
  }
#else
  constexpr int PS = SR;
  {
    const int rsub = lane / CPR, part = lane - rsub * CPR;
    if (lane < RPI * CPR) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int rec = w0 + RPI * q + rsub;
        if (rec < nobs) rec_cp16(sP + (size_t)rec * SR + part * EPC, Prec + (size_t)(o0 + rec) * REC + part * EPC);
      }
    }
  }
#endif
  int cam_idx = 0, lp = 0;
  if (t < nobs) { cam_idx = __ldg(a.view + o0 + t); lp = __ldg(a.point + o0 + t) - p0; }
  {
    const int r9 = lane / 9, c9 = lane - 9 * r9;
#pragma unroll
    for (int q = 0; q < 11; ++q) {
      const int rl = 3 * q + r9;                       // observation within the warp
      const int cam = __shfl_sync(FULL, cam_idx, rl & 31);
      if (lane < 27 && rl < 32 && w0 + rl < nobs) rec_cp1<T>(sD + 9 * (w0 + rl) + c9, dx_cam + 9 * (size_t)cam + c9);
    }
  }
  rec_commit();
  double acc_e = 0.0, acc_dx = 0.0, acc_jd = 0.0;
  int lo = 0, n = 0;
  T m0 = T(0), m1 = T(0);
  T R00 = T(1), R01 = T(0), R02 = T(0), R11 = T(1), R12v = T(0), R22 = T(1), c0 = T(0), c1 = T(0), c2 = T(0), G0 = T(0), G1 = T(0), G2 = T(0), pf = T(0);
  T X0 = T(0), X1 = T(0), X2 = T(0);
  Cam<T> cam;
  if (t < nobs) { m0 = __ldg(a.meas + 2 * (size_t)(o0 + t)); m1 = __ldg(a.meas + 2 * (size_t)(o0 + t) + 1); }
  const size_t gp = 3 * (size_t)(p0 + t);
  if (t < npts) {
    const int ps = __ldg(a.pt_start + p0 + t);
    lo = ps - o0; n = __ldg(a.pt_start + p0 + t + 1) - ps;
    const T* q = Ptrec + (size_t)(p0 + t) * PREC;
    T z_, z1_, z2_;
    load4(q, R00, R01, R02, R11);
    load4(q + 4, R12v, R22, c0, c1);
    load4(q + 8, c2, G0, G1, G2);
    load4(q + 12, pf, z_, z1_, z2_);
    X0 = __ldg(a.X + gp); X1 = __ldg(a.X + gp + 1); X2 = __ldg(a.X + gp + 2);
  }
  load_cam<T>(cams_test, cam_idx, cam);
  rec_wait<0>();
#ifdef BA_BULK_COPY
  if (wrec > 0) mbar_wait(&bars[t >> 5], 0);
#endif
  __syncwarp();
  if (t < nobs) {
    T r[REC], d[9];
    lds_rec<0, REC>(sP + (size_t)t * PS, r);
#pragma unroll
    for (int b = 0; b < 9; ++b) d[b] = sD[9 * t + b];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      T u = T(0);
#pragma unroll
      for (int b = 0; b < 9; ++b) u += r[9 * k + b] * d[b];
      su[k][t] = u;
    }
  }
  __syncthreads();
  if (t < npts) {
    T r0 = -c0, r1 = -c1, r2 = -c2;
    for (int i = 0; i < n; ++i) { r0 -= su[0][lo + i]; r1 -= su[1][lo + i]; r2 -= su[2][lo + i]; }
    const T z2 = r2 / R22;
    const T z1 = (r1 - R12v * z2) / R11;
    const T z0 = (r0 - R01 * z1 - R02 * z2) / R00;
    const int pm = (int)pf;
    T d[3];
    d[pm & 3] = z0; d[(pm >> 2) & 3] = z1; d[(pm >> 4) & 3] = z2;
    acc_dx += (double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    acc_jd += (double)(G0 * d[0] + G1 * d[1] + G2 * d[2]);
    dx_pt[gp] = d[0]; dx_pt[gp + 1] = d[1]; dx_pt[gp + 2] = d[2];
    const T x0 = X0 + d[0], x1 = X1 + d[1], x2 = X2 + d[2];
    X_test[gp] = x0; X_test[gp + 1] = x1; X_test[gp + 2] = x2;
    sx[0][t] = x0; sx[1][t] = x1; sx[2][t] = x2;
  }
  __syncthreads();
  if (t < nobs) {
    T e0, e1;
    obs_residual<T>(cam, sx[0][lp], sx[1][lp], sx[2][lp], m0, m1, a.tau2, e0, e1);
    acc_e += (double)(e0 * e0 + e1 * e1);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc_e += __shfl_down_sync(0xffffffffu, acc_e, off);
    acc_dx += __shfl_down_sync(0xffffffffu, acc_dx, off);
    acc_jd += __shfl_down_sync(0xffffffffu, acc_jd, off);
  }
  const int warp = t >> 5;
  if (lane == 0) { red[warp] = acc_e; red[TILE / 32 + warp] = acc_dx; red[2 * (TILE / 32) + warp] = acc_jd; }
  __syncthreads();
  if (t < 3) {
    double s = 0.0;
    for (int w = 0; w < TILE / 32; ++w) s += red[t * (TILE / 32) + w];
    partials[(size_t)t * ntiles + tile] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// MOREQR stage 2 (BacktrackLevMarqMore.h:293-348): per lambda trial only the 6x3 block [R0_j; sqrt(lambda) I3] of every
// point is re-triangularised (column-pivoted Householder, the conventions of the per-point QR above):
//   [R0; sqrt(lambda) I3] P' = Q' [R'; 0].   With Jp P0 = Q0 R0 from stage 1:  [Jp; sqrt(lambda) I3] P0 P' = Q [R'; 0],
// the observation rows of the thin Q being Q0_i Q'_11 (Q'_11 = top-left 3x3 of Q'). Hence, per observation,
//   Q1_i = Q0_i Q'_11,  R12'_i = Q1_i^T Jc_i,  c' = Q'_11^T c0,  M_i = I2 - Q1_i Q1_i^T,  w_i = e_i - Q1_i c',
// i.e. exactly the P / D / point records k_point_factor_warp writes for a trial, computed from the stage-1 J records
// without re-evaluating the Jacobian or re-factoring the (2 n_j + 3) x 3 block; the reduced-system kernels and the
// back-substitution then run unchanged (dx_j = P0 P' z: the two permutations are composed in the point record).
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void moreqr_inner_qr(const T (&R0)[6], const T sl, T (&R)[6], int& pm, T (&Q11)[3][3]) {
  const T tiny = (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f));
  T A[6][3] = {{R0[0], R0[1], R0[2]}, {T(0), R0[3], R0[4]}, {T(0), T(0), R0[5]}, {sl, T(0), T(0)}, {T(0), sl, T(0)}, {T(0), T(0), sl}};
  T tau[3];
  int pmv[3] = {0, 1, 2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T nn[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c >= k) {
#pragma unroll
        for (int r = 0; r < 6; ++r) if (r >= k) nn[c] += A[r][c] * A[r][c];
      }
    int best = k;
    T bestv = nn[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k && nn[c] > bestv) { best = c; bestv = nn[c]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) {
        const bool sw = best == c;
#pragma unroll
        for (int r = 0; r < 6; ++r) cswap(A[r][k], A[r][c], sw);
        const int t = pmv[k]; pmv[k] = sw ? pmv[c] : pmv[k]; pmv[c] = sw ? t : pmv[c];
      }
    }
    const T c0 = A[k][k];
    T tail2 = T(0);
#pragma unroll
    for (int r = 0; r < 6; ++r) if (r > k) tail2 += A[r][k] * A[r][k];
    const bool degenerate = tail2 <= tiny;
    T beta = tsqrt(c0 * c0 + tail2);
    if (c0 >= T(0)) beta = -beta;
    if (degenerate) beta = c0;
    const T inv = degenerate ? T(0) : T(1) / (c0 - beta);
    const T tk = degenerate ? T(0) : (beta - c0) / beta;
#pragma unroll
    for (int r = 0; r < 6; ++r) if (r > k) A[r][k] *= inv;
    A[k][k] = beta;
    tau[k] = tk;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) {
        T sdot = A[k][c];
#pragma unroll
        for (int r = 0; r < 6; ++r) if (r > k) sdot += A[r][k] * A[r][c];
        sdot *= tk;
        A[k][c] -= sdot;
#pragma unroll
        for (int r = 0; r < 6; ++r) if (r > k) A[r][c] -= sdot * A[r][k];
      }
    }
  }
  R[0] = A[0][0]; R[1] = A[0][1]; R[2] = A[0][2]; R[3] = A[1][1]; R[4] = A[1][2]; R[5] = A[2][2];
  pm = pmv[0] | (pmv[1] << 2) | (pmv[2] << 4);
  // top 3 rows of the thin Q: E = H0 H1 H2 [I3; 0], H_k = I - tau_k v_k v_k^T, v_k = (0.., 1 at row k, A[r][k] below)
  T E[6][3] = {{T(1), T(0), T(0)}, {T(0), T(1), T(0)}, {T(0), T(0), T(1)}, {T(0), T(0), T(0)}, {T(0), T(0), T(0)}, {T(0), T(0), T(0)}};
#pragma unroll
  for (int k = 2; k >= 0; --k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T sdot = E[k][c];
#pragma unroll
      for (int r = 0; r < 6; ++r) if (r > k) sdot += A[r][k] * E[r][c];
      sdot *= tau[k];
      E[k][c] -= sdot;
#pragma unroll
      for (int r = 0; r < 6; ++r) if (r > k) E[r][c] -= sdot * A[r][k];
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Q11[r][c] = E[r][c];
}

template <class T>
__global__ void __launch_bounds__(TILE, 4) k_moreqr_stage2(const T lambda, int nunits, const int* __restrict__ unit_obs, const int* __restrict__ seg,
                                                           const int* __restrict__ slot, const int* __restrict__ point, const T* __restrict__ Jrec,
                                                           const T* __restrict__ Ptrec0, T* __restrict__ Prec, T* __restrict__ Drec, T* __restrict__ Ptrec) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int EPC = RecGeom<T>::EPC, CPR = RecGeom<T>::CPR, SR = RecGeom<T>::SREC;
  constexpr int RPI = 32 / CPR, NQ = 32 / RPI;
  const int lane = threadIdx.x & 31, u = blockIdx.x * (TILE / 32) + (threadIdx.x >> 5);
  if (u >= nunits) return;
  extern __shared__ __align__(16) unsigned char mq_smem_raw[];
  T* const my = reinterpret_cast<T*>(mq_smem_raw) + (size_t)(threadIdx.x >> 5) * 32 * SR;
  const int2 uo = __ldg(reinterpret_cast<const int2*>(unit_obs) + u);
  const int o0 = uo.x, un = uo.y;
  const bool act = lane < un;
  const int o = o0 + (act ? lane : 0);
  const int rsub = lane / CPR, part = lane - rsub * CPR;
  const bool cpl = lane < RPI * CPR;
  // the unit's J records are contiguous: staged with coalesced cp.async (lane = (record, 16-byte chunk))
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int r = RPI * q + rsub;
    if (cpl && r < un) rec_cp16(my + r * SR + part * EPC, Jrec + (size_t)(o0 + r) * REC + part * EPC);
  }
  rec_commit();
  const int pj = __ldg(point + o), sg = __ldg(seg + o);
  const size_t sl = (size_t)__ldg(slot + o);
  const int i = act ? (sg & 0xff) : 0;
  T R0[6], c0[3], G[3], pf, z0_, z1_, z2_;
  {
    const T* q = Ptrec0 + (size_t)pj * PREC;
    load4(q, R0[0], R0[1], R0[2], R0[3]);
    load4(q + 4, R0[4], R0[5], c0[0], c0[1]);
    load4(q + 8, c0[2], G[0], G[1], G[2]);
    load4(q + 12, pf, z0_, z1_, z2_);
  }
  T R[6], Q11[3][3]; int pm1 = 0;
  moreqr_inner_qr<T>(R0, tsqrt(lambda), R, pm1, Q11);
  const int pm0 = (int)pf;
  // dx_j (original order) = P0 P' z: column a of R' is column pm1[a] of R0, i.e. original column pm0[pm1[a]]
  int pm = 0;
#pragma unroll
  for (int a2 = 0; a2 < 3; ++a2) { const int m = (pm1 >> (2 * a2)) & 3; pm |= ((pm0 >> (2 * m)) & 3) << (2 * a2); }
  T cq[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) cq[k] = Q11[0][k] * c0[0] + Q11[1][k] * c0[1] + Q11[2][k] * c0[2];
  rec_wait<0>();
  __syncwarp();
  T jr[REC];
  lds_rec<0, REC>(my + lane * SR, jr);
  __syncwarp();
  T x[2][3];
#pragma unroll
  for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
    for (int k = 0; k < 3; ++k) x[a2][k] = jr[18 + 3 * a2] * Q11[0][k] + jr[18 + 3 * a2 + 1] * Q11[1][k] + jr[18 + 3 * a2 + 2] * Q11[2][k];
  const T e0 = jr[24], e1 = jr[25];
  T rec[REC];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int b = 0; b < 9; ++b) rec[9 * k + b] = x[0][k] * jr[b] + x[1][k] * jr[9 + b];
  rec[27] = T(0);
  store_rec(my + lane * SR, rec);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int r = RPI * q + rsub;
    if (cpl && r < un)
      *reinterpret_cast<int4*>(Prec + (size_t)(o0 + r) * REC + part * EPC) = *reinterpret_cast<const int4*>(my + r * SR + part * EPC);
  }
  __syncwarp();
#pragma unroll
  for (int b = 0; b < 18; ++b) rec[b] = jr[b];
  fill_drec_tail<T>(rec, x[0][0], x[0][1], x[0][2], x[1][0], x[1][1], x[1][2], cq[0], cq[1], cq[2], e0, e1);
  store_rec(my + lane * SR, rec);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int r = RPI * q + rsub;
    const size_t slr = (size_t)__shfl_sync(FULL, (int)sl, r & 31);
    if (cpl && r < un)
      *reinterpret_cast<int4*>(Drec + slr * REC + part * EPC) = *reinterpret_cast<const int4*>(my + r * SR + part * EPC);
  }
  if (act && i == 0) {
    T* q = Ptrec + (size_t)pj * PREC;
    store4(q, R[0], R[1], R[2], R[3]);
    store4(q + 4, R[4], R[5], cq[0], cq[1]);
    store4(q + 8, cq[2], G[0], G[1], G[2]);
    store4(q + 12, (T)pm, T(0), T(0), T(0));
  }
}

// ---------------------------------------------------------------------------------------------
// Right block of QRKIT / MOREQR (DenseBlockedThinQR of J2bot, BAFunctor.h:101,111; call sites More.h:288,328). The tall
// J2bot ((2K+9N) x 9N dense: 1.3 TB at BASELINE config 5) cannot be materialised, and a Householder QR of the SQUARE
// S = J2bot^T J2bot has the accuracy of the normal equations at 4-10x the cost of an LDL^T. What a QR of J2bot buys is a
// solution whose error does not carry cond(S) = cond(J2bot)^2 from FORMING S; the same is obtained with the corrected
// semi-normal equations (Bjorck 1987): y0 from the factor of S, then the residual of the LEAST-SQUARES problem taken
// through J2bot itself, never through S,
//     rho = d - J2bot y0,   J2bot^T rho = sum_i Jc_i^T u_i - lambda y0,   u = (I - Q1 Q1^T)(e - Jc y0)  (per point),
// one more solve S delta = J2bot^T rho, y = y0 + delta. J2bot is never formed: Jc_i, Q1_i, e_i sit in the D records.
//   k_csne_point (point-major, warp units, lane = observation; k_csne_point_long: one warp per point with more than 32
//   observations): u_i into the observation's camera-major slot;  k_csne_cam (one CTA per camera, fixed-order sums):
//   r_a = sum Jc_i^T u_i + lambda dx_cam_a  (dx_cam = -y0).
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void csne_load(const T* __restrict__ Drec, const size_t sl, const T* __restrict__ dx_cam, const int cam, T (&q)[2][3], T& et0, T& et1) {
  T r[REC];
#pragma unroll
  for (int c = 0; c < REC; c += 4) load4(Drec + sl * REC + c, r[c], r[c + 1], r[c + 2], r[c + 3]);
  et0 = r[26]; et1 = r[27];
#pragma unroll
  for (int b = 0; b < 9; ++b) { const T d = __ldg(dx_cam + 9 * (size_t)cam + b); et0 += r[b] * d; et1 += r[9 + b] * d; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { q[0][k] = r[18 + k]; q[1][k] = r[21 + k]; }
}

template <class T>
__global__ void __launch_bounds__(TILE) k_csne_point(int nunits, const int* __restrict__ unit_obs, const int* __restrict__ seg, const int* __restrict__ slot,
                                                     const int* __restrict__ view, const T* __restrict__ Drec, const T* __restrict__ dx_cam, T* __restrict__ uarr) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, u = blockIdx.x * (TILE / 32) + (threadIdx.x >> 5);
  if (u >= nunits) return;
  const int2 uo = __ldg(reinterpret_cast<const int2*>(unit_obs) + u);
  const int o0 = uo.x, un = uo.y;
  const bool act = lane < un;
  const int o = o0 + (act ? lane : 0);
  const int cam = __ldg(view + o), sg = __ldg(seg + o);
  const size_t sl = (size_t)__ldg(slot + o);
  const int i = act ? (sg & 0xff) : 0, n = act ? (sg >> 8) : 1;
  const int s0 = lane - i;
  T q[2][3], et0, et1;
  csne_load<T>(Drec, sl, dx_cam, cam, q, et0, et1);
  if (!act) { et0 = et1 = T(0); }
  const int nmax = __reduce_max_sync(FULL, n);
  T t[3] = {q[0][0] * et0 + q[1][0] * et1, q[0][1] * et0 + q[1][1] * et1, q[0][2] * et0 + q[1][2] * et1};
  seg_sum<T, 3>(t, s0, n, nmax, i);
  if (act) {
    uarr[2 * sl] = et0 - (q[0][0] * t[0] + q[0][1] * t[1] + q[0][2] * t[2]);
    uarr[2 * sl + 1] = et1 - (q[1][0] * t[0] + q[1][1] * t[1] + q[1][2] * t[2]);
  }
}

// points with more than 32 observations: one warp per point, two sweeps over its observations (fixed order)
template <class T>
__global__ void __launch_bounds__(TILE) k_csne_point_long(int nlong, const int* __restrict__ long_pt, const int* __restrict__ pt_start, const int* __restrict__ slot,
                                                          const int* __restrict__ view, const T* __restrict__ Drec, const T* __restrict__ dx_cam, T* __restrict__ uarr) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = blockIdx.x * (TILE / 32) + (threadIdx.x >> 5);
  if (w >= nlong) return;
  const int pj = __ldg(long_pt + w), o0 = __ldg(pt_start + pj), n = __ldg(pt_start + pj + 1) - o0;
  T t[3] = {T(0), T(0), T(0)};
  for (int base = 0; base < n; base += 32) {
    const int ii = base + lane;
    T c3[3] = {T(0), T(0), T(0)};
    if (ii < n) {
      T q[2][3], et0, et1;
      csne_load<T>(Drec, (size_t)__ldg(slot + o0 + ii), dx_cam, __ldg(view + o0 + ii), q, et0, et1);
#pragma unroll
      for (int k = 0; k < 3; ++k) c3[k] = q[0][k] * et0 + q[1][k] * et1;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) c3[k] += __shfl_down_sync(FULL, c3[k], off);
      t[k] += __shfl_sync(FULL, c3[k], 0);
    }
  }
  for (int base = 0; base < n; base += 32) {
    const int ii = base + lane;
    if (ii < n) {
      T q[2][3], et0, et1;
      const size_t sl = (size_t)__ldg(slot + o0 + ii);
      csne_load<T>(Drec, sl, dx_cam, __ldg(view + o0 + ii), q, et0, et1);
      uarr[2 * sl] = et0 - (q[0][0] * t[0] + q[0][1] * t[1] + q[0][2] * t[2]);
      uarr[2 * sl + 1] = et1 - (q[1][0] * t[0] + q[1][1] * t[1] + q[1][2] * t[2]);
    }
  }
}

template <class T>
__global__ void __launch_bounds__(128) k_csne_cam(const int* __restrict__ cam_start, const T* __restrict__ Drec, const T* __restrict__ uarr,
                                                  const T* __restrict__ dx_cam, const T lambda_diag, T* __restrict__ rout /* 9 per camera */) {
  __shared__ T part[4][9];
  const int cam = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T s[9];
#pragma unroll
  for (int b = 0; b < 9; ++b) s[b] = T(0);
  for (int q = __ldg(cam_start + cam) + threadIdx.x; q < __ldg(cam_start + cam + 1); q += 128) {
    const T* rec = Drec + (size_t)q * REC;
    T jc[20];
#pragma unroll
    for (int c = 0; c < 20; c += 4) load4(rec + c, jc[c], jc[c + 1], jc[c + 2], jc[c + 3]);
    const T u0 = __ldg(uarr + 2 * (size_t)q), u1 = __ldg(uarr + 2 * (size_t)q + 1);
#pragma unroll
    for (int b = 0; b < 9; ++b) s[b] += jc[b] * u0 + jc[9 + b] * u1;
  }
#pragma unroll
  for (int b = 0; b < 9; ++b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s[b] += __shfl_down_sync(0xffffffffu, s[b], off);
    if (lane == 0) part[warp][b] = s[b];
  }
  __syncthreads();
  if (threadIdx.x < 9)
    rout[9 * (size_t)cam + threadIdx.x] = (((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x]) + lambda_diag * dx_cam[9 * (size_t)cam + threadIdx.x];
}

// dx_cam += delta
template <class T>
__global__ void k_axpy1(int n, const T* __restrict__ x, T* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += x[i];
}

// ---------------------------------------------------------------------------------------------
// Points with more than TILE observations (long tracks in real BAL files): one CTA per point. The Jp rows and the
// residual of all observations sit in dynamic shared memory (8 scalars per observation), the 3-column
// factorisation is the same routine as in the tile kernel (one thread, serial: such points are rare), the camera
// Jacobian is re-evaluated for the records. Same records, same conventions as k_point_factor_warp.
// ---------------------------------------------------------------------------------------------
constexpr int BIG_THREADS = 256;
constexpr int BIG_POINT_MAX = 3000;  // 8 scalars per observation in shared memory (192 KB in double)
template <class T> constexpr size_t big_point_smem_bytes(int nmax) { return (size_t)8 * nmax * sizeof(T); }

// ---- block-parallel versions of point_householder / point_normal for ONE point per CTA (long tracks: up to 3000
// observations). Thread t owns the observations t, t + BIG_THREADS, ...; every sum over the point's rows (column norms,
// reflector dots, Q1^T e, Jp^T e, the 3x3 normal equations) is a fixed-order block reduction whose result every thread
// receives, so the 3x3 R, the reflector scalars and the sqrt(lambda) I3 rows are simply replicated in all threads (as in
// the warp-segmented kernel). Round 1 ran the factorisation of such a point on one thread.
template <class T, int NV>
__device__ __forceinline__ void big_block_sum(T (&v)[NV], T* red /* [BIG_THREADS / 32][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
  }
  __syncthreads();                       // previous use of `red` is over
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) red[warp * NV + q] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < BIG_THREADS / 32; ++w) s += red[w * NV + q];
    v[q] = s;
  }
}

template <class T>
__device__ void big_point_householder(BigPointStore<T>& st, const int n, const T sl, T* red) {
  const int t = threadIdx.x, S = st.st;
  const T tiny = (sizeof(T) == 8 ? T(2.2250738585072014e-308) : T(1.17549435e-38f));
  T L[3][3] = {{sl, T(0), T(0)}, {T(0), sl, T(0)}, {T(0), T(0), sl}};
  T tau[3];
  int pmv[3] = {0, 1, 2};
  auto Q = [&](int a, int b, int i) -> T& { return st.Q[(size_t)(3 * a + b) * S + i]; };
  {  // G = Jp^T e in the original column order
    T g[3] = {T(0), T(0), T(0)};
    for (int i = t; i < n; i += BIG_THREADS) {
      const T e0 = st.E[i], e1 = st.E[S + i];
#pragma unroll
      for (int b = 0; b < 3; ++b) g[b] += Q(0, b, i) * e0 + Q(1, b, i) * e1;
    }
    big_block_sum<T, 3>(g, red);
    if (t == 0) { st.G[0] = g[0]; st.G[1] = g[1]; st.G[2] = g[2]; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // squared norms of the remaining columns over rows >= k (row rho = 2 i + a)
    T nn[3] = {T(0), T(0), T(0)};
    for (int i = t; i < n; i += BIG_THREADS) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (2 * i + a >= k) {
#pragma unroll
          for (int c = 0; c < 3; ++c) if (c >= k) { const T v = Q(a, c, i); nn[c] += v * v; }
        }
      }
    }
    big_block_sum<T, 3>(nn, red);
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c >= k) nn[c] += L[0][c] * L[0][c] + L[1][c] * L[1][c] + L[2][c] * L[2][c];
    int best = k;
    T bestv = nn[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k && nn[c] > bestv) { best = c; bestv = nn[c]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) {
        const bool sw = best == c;
        if (sw) {
          for (int i = t; i < n; i += BIG_THREADS) {
#pragma unroll
            for (int a = 0; a < 2; ++a) { const T x = Q(a, k, i); Q(a, k, i) = Q(a, c, i); Q(a, c, i) = x; }
          }
        }
        cswap(L[0][k], L[0][c], sw); cswap(L[1][k], L[1][c], sw); cswap(L[2][k], L[2][c], sw);
        const int x = pmv[k]; pmv[k] = sw ? pmv[c] : pmv[k]; pmv[c] = sw ? x : pmv[c];
      }
    }
    __syncthreads();
    // row k lives in observation k >> 1, slot k & 1
    const T c0 = Q(k & 1, k, k >> 1);
    T akc[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k) akc[c] = Q(k & 1, c, k >> 1);
    T sums[3] = {T(0), T(0), T(0)};   // tail2, dot with column k+1, dot with column k+2 (rows > k)
    for (int i = t; i < n; i += BIG_THREADS) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (2 * i + a > k) {
          const T v = Q(a, k, i);
          sums[0] += v * v;
#pragma unroll
          for (int c = 0; c < 3; ++c) if (c > k) sums[c - k] += v * Q(a, c, i);
        }
      }
    }
    big_block_sum<T, 3>(sums, red);
    const T tail2 = sums[0] + L[0][k] * L[0][k] + L[1][k] * L[1][k] + L[2][k] * L[2][k];
    const bool degenerate = tail2 <= tiny;
    T beta = tsqrt(c0 * c0 + tail2);
    if (c0 >= T(0)) beta = -beta;
    if (degenerate) beta = c0;
    const T inv = degenerate ? T(0) : T(1) / (c0 - beta);
    const T tk = degenerate ? T(0) : (beta - c0) / beta;
    tau[k] = tk;
    // the dots were taken with the UN-scaled column k: v_r = A[r][k] inv
    T sd[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c > k) sd[c] = (akc[c] + inv * sums[c - k] + inv * (L[0][k] * L[0][c] + L[1][k] * L[1][c] + L[2][k] * L[2][c])) * tk;
    for (int i = t; i < n; i += BIG_THREADS) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int rho = 2 * i + a;
        if (rho > k) {
          const T v = Q(a, k, i) * inv;
          Q(a, k, i) = v;
#pragma unroll
          for (int c = 0; c < 3; ++c) if (c > k) Q(a, c, i) -= sd[c] * v;
        } else if (rho == k) {
          Q(a, k, i) = beta;
#pragma unroll
          for (int c = 0; c < 3; ++c) if (c > k) Q(a, c, i) -= sd[c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) { L[0][c] -= sd[c] * (L[0][k] * inv); L[1][c] -= sd[c] * (L[1][k] * inv); L[2][c] -= sd[c] * (L[2][k] * inv); }
    }
    L[0][k] *= inv; L[1][k] *= inv; L[2][k] *= inv;
    __syncthreads();
  }
  if (t == 0) {  // R sits in rows 0..2 (observation 0: rows 0, 1; observation 1: row 2), already in pivoted column order
    st.Rm[0] = Q(0, 0, 0); st.Rm[1] = Q(0, 1, 0); st.Rm[2] = Q(0, 2, 0);
    st.Rm[3] = Q(1, 1, 0); st.Rm[4] = Q(1, 2, 0); st.Rm[5] = Q(0, 2, 1);
    st.perm[0] = pmv[0] | (pmv[1] << 2) | (pmv[2] << 4);
  }
  __syncthreads();
  // form the thin Q1 in place (dorg2r): k = 2, 1, 0
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const T tk = tau[k];
    T akc[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k) akc[c] = Q(k & 1, c, k >> 1);
    T dots[2] = {T(0), T(0)};
    for (int i = t; i < n; i += BIG_THREADS) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (2 * i + a > k) {
          const T v = Q(a, k, i);
#pragma unroll
          for (int c = 0; c < 3; ++c) if (c > k) dots[c - k - 1] += v * Q(a, c, i);
        }
      }
    }
    big_block_sum<T, 2>(dots, red);
    T sd[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c > k) sd[c] = (akc[c] + dots[c - k - 1] + L[0][k] * L[0][c] + L[1][k] * L[1][c] + L[2][k] * L[2][c]) * tk;
    for (int i = t; i < n; i += BIG_THREADS) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int rho = 2 * i + a;
        if (rho > k) {
          const T v = Q(a, k, i);
#pragma unroll
          for (int c = 0; c < 3; ++c) if (c > k) Q(a, c, i) -= sd[c] * v;
          Q(a, k, i) = -tk * v;
        } else if (rho == k) {
#pragma unroll
          for (int c = 0; c < 3; ++c) if (c > k) Q(a, c, i) -= sd[c];
          Q(a, k, i) = T(1) - tk;
        } else {
          Q(a, k, i) = T(0);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c > k) { L[0][c] -= sd[c] * L[0][k]; L[1][c] -= sd[c] * L[1][k]; L[2][c] -= sd[c] * L[2][k]; }
    }
    L[0][k] *= -tk; L[1][k] *= -tk; L[2][k] *= -tk;
    __syncthreads();
  }
  T cq[3] = {T(0), T(0), T(0)};
  for (int i = t; i < n; i += BIG_THREADS) {
    const T e0 = st.E[i], e1 = st.E[S + i];
#pragma unroll
    for (int b = 0; b < 3; ++b) cq[b] += Q(0, b, i) * e0 + Q(1, b, i) * e1;
  }
  big_block_sum<T, 3>(cq, red);
  if (t == 0) { st.C[0] = cq[0]; st.C[1] = cq[1]; st.C[2] = cq[2]; }
  __syncthreads();
}

// normal-equation point factor (CHOLESKY variant), one point per CTA
template <class T>
__device__ void big_point_normal(BigPointStore<T>& st, const int n, const T lambda, T* red) {
  const int t = threadIdx.x, S = st.st;
  auto Q = [&](int a, int b, int i) -> T& { return st.Q[(size_t)(3 * a + b) * S + i]; };
  T v[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, g[3] = {T(0), T(0), T(0)};
  for (int i = t; i < n; i += BIG_THREADS) {
    const T e0 = st.E[i], e1 = st.E[S + i];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const T x0 = Q(a, 0, i), x1 = Q(a, 1, i), x2 = Q(a, 2, i), e = a ? e1 : e0;
      v[0] += x0 * x0; v[1] += x1 * x0; v[2] += x1 * x1; v[3] += x2 * x0; v[4] += x2 * x1; v[5] += x2 * x2;
      g[0] += x0 * e; g[1] += x1 * e; g[2] += x2 * e;
    }
  }
  big_block_sum<T, 6>(v, red);
  big_block_sum<T, 3>(g, red);
  const T v00 = v[0] + lambda, v10 = v[1], v11 = v[2] + lambda, v20 = v[3], v21 = v[4], v22 = v[5] + lambda;
  const T d0 = v00, l10 = v10 / d0, l20 = v20 / d0;
  const T d1 = tmax(v11 - l10 * l10 * d0, lambda), l21 = (v21 - l20 * l10 * d0) / d1;   // pivot floor: see point_normal
  const T d2 = tmax(v22 - l20 * l20 * d0 - l21 * l21 * d1, lambda);
  const T q0 = tsqrt(d0), q1 = tsqrt(d1), q2 = tsqrt(d2);
  const T R[6] = {q0, q0 * l10, q0 * l20, q1, q1 * l21, q2};
  const T i00 = T(1) / R[0], i11 = T(1) / R[3], i22 = T(1) / R[5];
  T cq[3] = {T(0), T(0), T(0)};
  for (int i = t; i < n; i += BIG_THREADS) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const T y0 = Q(a, 0, i) * i00;
      const T y1 = (Q(a, 1, i) - y0 * R[1]) * i11;
      const T y2 = (Q(a, 2, i) - y0 * R[2] - y1 * R[4]) * i22;
      Q(a, 0, i) = y0; Q(a, 1, i) = y1; Q(a, 2, i) = y2;
      const T e = st.E[(size_t)a * S + i];
      cq[0] += y0 * e; cq[1] += y1 * e; cq[2] += y2 * e;
    }
  }
  big_block_sum<T, 3>(cq, red);
  if (t == 0) {
#pragma unroll
    for (int q = 0; q < 6; ++q) st.Rm[q] = R[q];
    st.perm[0] = 0 | (1 << 2) | (2 << 4);
    st.G[0] = g[0]; st.G[1] = g[1]; st.G[2] = g[2];
    st.C[0] = cq[0]; st.C[1] = cq[1]; st.C[2] = cq[2];
  }
  __syncthreads();
}

template <class T>
__global__ void __launch_bounds__(BIG_THREADS) k_point_factor_big(TileArgs<T> a, const int* __restrict__ huge_pt, const int* __restrict__ slot,
                                                                  T* __restrict__ Prec, T* __restrict__ Drec, T* __restrict__ Ptrec) {
  extern __shared__ __align__(16) unsigned char big_smem_raw[];
  __shared__ BigPointStore<T> st;
  const int t = threadIdx.x, pj = __ldg(huge_pt + blockIdx.x);
  const int o0 = __ldg(a.pt_start + pj), n = __ldg(a.pt_start + pj + 1) - o0;
  if (t == 0) { st.Q = reinterpret_cast<T*>(big_smem_raw); st.E = st.Q + 6 * (size_t)n; st.st = n; }
  __syncthreads();
  const T X0 = __ldg(a.X + 3 * (size_t)pj), X1 = __ldg(a.X + 3 * (size_t)pj + 1), X2 = __ldg(a.X + 3 * (size_t)pj + 2);
  for (int i = t; i < n; i += BIG_THREADS) {
    const int o = o0 + i;
    Cam<T> c; load_cam<T>(a.cams, __ldg(a.view + o), c);
    T e0, e1, jc[18], jp[6];
    obs_jacobian<T>(c, X0, X1, X2, __ldg(a.meas + 2 * (size_t)o), __ldg(a.meas + 2 * (size_t)o + 1), a.tau2, e0, e1, jc, jp);
    st.E[i] = e0; st.E[n + i] = e1;
#pragma unroll
    for (int k = 0; k < 6; ++k) st.Q[(size_t)k * n + i] = jp[k];
  }
  __syncthreads();
  __shared__ T bred[(BIG_THREADS / 32) * 6];
  if (a.factor == PF_HOUSEHOLDER && n >= 2) big_point_householder<T>(st, n, tsqrt(a.lambda), bred);
  else big_point_normal<T>(st, n, a.lambda, bred);
  __syncthreads();
  const T c0 = st.C[0], c1 = st.C[1], c2 = st.C[2];
  for (int i = t; i < n; i += BIG_THREADS) {
    const int o = o0 + i;
    Cam<T> c; load_cam<T>(a.cams, __ldg(a.view + o), c);
    T e0, e1, jc[18], jp[6];
    obs_jacobian<T>(c, X0, X1, X2, __ldg(a.meas + 2 * (size_t)o), __ldg(a.meas + 2 * (size_t)o + 1), a.tau2, e0, e1, jc, jp);
    T q[2][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { q[0][k] = st.Q[(size_t)k * n + i]; q[1][k] = st.Q[(size_t)(3 + k) * n + i]; }
    T rec[REC];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int b = 0; b < 9; ++b) rec[9 * k + b] = q[0][k] * jc[b] + q[1][k] * jc[9 + b];
    rec[27] = T(0);
    store_rec(Prec + (size_t)o * REC, rec);
#pragma unroll
    for (int b = 0; b < 18; ++b) rec[b] = jc[b];
    fill_drec_tail<T>(rec, q[0][0], q[0][1], q[0][2], q[1][0], q[1][1], q[1][2], c0, c1, c2, e0, e1);
    store_rec(Drec + (size_t)__ldg(slot + o) * REC, rec);
  }
  if (t == 0) {
    T* q = Ptrec + (size_t)pj * PREC;
    store4(q, st.Rm[0], st.Rm[1], st.Rm[2], st.Rm[3]);
    store4(q + 4, st.Rm[4], st.Rm[5], st.C[0], st.C[1]);
    store4(q + 8, st.C[2], st.G[0], st.G[1], st.G[2]);
    store4(q + 12, (T)st.perm[0], T(0), T(0), T(0));
  }
}

// fixed-order block sum of three doubles (BIG_THREADS threads); result valid in thread 0
__device__ __forceinline__ void big_block_sum3(double& a0, double& a1, double& a2, double* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a0 += __shfl_down_sync(0xffffffffu, a0, off);
    a1 += __shfl_down_sync(0xffffffffu, a1, off);
    a2 += __shfl_down_sync(0xffffffffu, a2, off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { red[warp] = a0; red[8 + warp] = a1; red[16 + warp] = a2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a0 = a1 = a2 = 0.0;
    for (int w = 0; w < BIG_THREADS / 32; ++w) { a0 += red[w]; a1 += red[8 + w]; a2 += red[16 + w]; }
  }
}

// back-substitution + update + test energy of one point with more than TILE observations (k_backsub_eval's job);
// partial sums go to tile slot `tile0 + blockIdx.x` of the same partials array.
template <class T>
__global__ void __launch_bounds__(BIG_THREADS) k_backsub_big(TileArgs<T> a, const int* __restrict__ huge_pt, const T* __restrict__ Prec,
                                                             const T* __restrict__ Ptrec, const T* __restrict__ dx_cam, const T* __restrict__ cams_test,
                                                             T* __restrict__ dx_pt, T* __restrict__ X_test, double* __restrict__ partials, int tile0,
                                                             int ntiles) {
  __shared__ double red[24];
  __shared__ T sx[3];
  const int t = threadIdx.x, pj = __ldg(huge_pt + blockIdx.x);
  const int o0 = __ldg(a.pt_start + pj), n = __ldg(a.pt_start + pj + 1) - o0;
  double u0 = 0.0, u1 = 0.0, u2 = 0.0;
  for (int i = t; i < n; i += BIG_THREADS) {
    const int o = o0 + i;
    const T* pr = Prec + (size_t)o * REC;
    const T* d = dx_cam + 9 * (size_t)__ldg(a.view + o);
    T dd[9];
#pragma unroll
    for (int b = 0; b < 9; ++b) dd[b] = __ldg(d + b);
    T v[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int b = 0; b < 9; ++b) v[k] += __ldg(pr + 9 * k + b) * dd[b];
    u0 += (double)v[0]; u1 += (double)v[1]; u2 += (double)v[2];
  }
  big_block_sum3(u0, u1, u2, red);
  double acc_e = 0.0, acc_dx = 0.0, acc_jd = 0.0;
  if (t == 0) {
    const T* q = Ptrec + (size_t)pj * PREC;
    T R00, R01, R02, R11, R12v, R22, c0, c1, c2, G0, G1, G2, pf, z_, z1_, z2_;
    load4(q, R00, R01, R02, R11);
    load4(q + 4, R12v, R22, c0, c1);
    load4(q + 8, c2, G0, G1, G2);
    load4(q + 12, pf, z_, z1_, z2_);
    const T r0 = -c0 - (T)u0, r1 = -c1 - (T)u1, r2 = -c2 - (T)u2;
    const T z2 = r2 / R22;
    const T z1 = (r1 - R12v * z2) / R11;
    const T z0 = (r0 - R01 * z1 - R02 * z2) / R00;
    const int pm = (int)pf;
    T d[3];
    d[pm & 3] = z0; d[(pm >> 2) & 3] = z1; d[(pm >> 4) & 3] = z2;
    const size_t gp = 3 * (size_t)pj;
    acc_dx = (double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    acc_jd = (double)(G0 * d[0] + G1 * d[1] + G2 * d[2]);
    dx_pt[gp] = d[0]; dx_pt[gp + 1] = d[1]; dx_pt[gp + 2] = d[2];
    const T x0 = __ldg(a.X + gp) + d[0], x1 = __ldg(a.X + gp + 1) + d[1], x2 = __ldg(a.X + gp + 2) + d[2];
    X_test[gp] = x0; X_test[gp + 1] = x1; X_test[gp + 2] = x2;
    sx[0] = x0; sx[1] = x1; sx[2] = x2;
  }
  __syncthreads();
  for (int i = t; i < n; i += BIG_THREADS) {
    const int o = o0 + i;
    Cam<T> c; load_cam<T>(cams_test, __ldg(a.view + o), c);
    T e0, e1;
    obs_residual<T>(c, sx[0], sx[1], sx[2], __ldg(a.meas + 2 * (size_t)o), __ldg(a.meas + 2 * (size_t)o + 1), a.tau2, e0, e1);
    acc_e += (double)(e0 * e0 + e1 * e1);
  }
  big_block_sum3(acc_e, acc_dx, acc_jd, red);
  if (t == 0) {
    const size_t idx = (size_t)tile0 + blockIdx.x;
    partials[idx] = acc_e; partials[(size_t)ntiles + idx] = acc_dx; partials[2 * (size_t)ntiles + idx] = acc_jd;
  }
}

}  // namespace ba

// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU restatement ("oracle") of the Levenberg-Marquardt inner step of
// jasvob/BundleAdjustment_Benchmarks. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this. The product (csrc/) never links it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or logs for this path, and its
// sparse-QR arithmetic lives in an un-vendored, un-pinned private Eigen fork ("eigen_sparse_qr" /
// QRkit: BlockDiagonalSparseQR, BlockAngularSparseQR[Partial], DenseBlockedThinQR,
// Eigen::BacktrackLevMarq) plus stock Eigen ColPivHouseholderQR / SimplicialLDLT (version
// unpinned). The reference cannot be compiled here (no Eigen, no SuiteSparse, MSVC-only sources).
// What this file follows, function by function:
//   model / Jacobian / update : src/Optimization/BAFunctor.h:126-342, src/DistortionFunction.cpp:14-51,
//                               src/CameraMatrix.cpp:207-209,259-261, src/MathUtils.h:13-21,66-82
//   LM control flow + constants: src/Eigen_ext/BacktrackLevMarqQRChol.h:131-160,204-436 (QRCHOL, and
//                               the structural template for the absent QRKIT loop),
//                               BacktrackLevMarqMore.h:204-425 (MOREQR), BacktrackLevMarqCholesky.h:190-361
//   block shapes              : src/Optimization/BAFunctor.cpp:64-78; row permutation QRChol.h:291-315
// The missing solver classes are restated from their published algorithms: per-point dense
// column-pivoted Householder QR (Eigen ColPivHouseholderQR conventions), Q^T applied to the camera
// columns and residual, the reduced camera system J2bot^T J2bot formed EXPLICITLY from J2bot (as
// QRChol.h:339 does), un-pivoted LDL^T (SimplicialLDLT arithmetic, natural order) or Householder QR
// for the right block, upper-triangular back-substitution.
// It is pinned instead by: finite-difference Jacobian checks, an independent NumPy restatement
// (oracle/ba_oracle_np.py) with an extended-precision dense solve, and the survey-time anchors in
// tests/golden/ (see DESIGN.md "Oracle").
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace bao {

enum Variant { QRKIT = 0, QRCHOL = 1, MOREQR = 2, CHOLESKY = 3 };

struct TrialRecord {  // one row of the reference's iteration table (QRChol.h:74-93) + |dx|
  int iter;
  int accepted;
  double energy;       // m_energy before the step (what the table prints, quirk Q4)
  double energy_test;  // energy at the test point
  double rho;
  double lambda_used;  // lambda this trial was solved with
  double lambda_next;  // lambda after the accept/reject update (what the table prints)
  double dx_norm;
};

enum Status { NotStarted = -2, Running = -1, Success = 0, ExceededLambdaMax = 1,
              TooManyFunctionEvaluation = 2, MaxItersReached = 3 };  // QRChol.h:39-46

template <class S>
struct Oracle {
  int N = 0, M = 0, K = 0;
  std::vector<int> view, point, pt_start;  // observations sorted by point; CSR offsets
  std::vector<S> meas;                     // 2K
  S tau = S(0.5);
  // state (BAFunctor.h:39-51 InputType): R row-major 3x3, T, f(=K00, negative), k1, k2; X
  std::vector<S> R, T, f, k1, k2, X;
  // linearisation at x
  std::vector<S> res;     // 2K
  std::vector<S> Jc, Jp;  // K*18 (2x9 row-major), K*6 (2x3 row-major)
  std::vector<S> JtRes;   // 3M+9N  (= -J^T r, QRChol.h:267)
  // reduced system of the last step (band storage, lower): Sb[i*(kd+1) + (j-i+kd)]
  int bw = 0, kd = 0;
  std::vector<S> Sb, g;
  std::vector<S> S_last, g_last;  // copy kept for tests (before factorisation)
  // MOREQR stage-1 storage
  std::vector<S> m_R, m_c, m_R12, m_S0, m_g0;
  std::vector<int> m_perm;
  bool tall_qr = false;  // QRKIT only: reference-faithful QR of the tall J2bot (small problems)
  // Timing legs only (bench.py cpu_baseline "all cores"): > 1 runs the per-observation / per-point loops and the
  // reduced solve with OpenMP. The parity tests use the default 1 (the reference is single-threaded and the
  // summation order of the single-threaded code is the one the golden anchors were taken with).
  int nthreads = 1;
  S eps_psi = S(1e-15);  // BAFunctor.h:159

  int nparams() const { return 3 * M + 9 * N; }

  void init(int N_, int M_, int K_, const int* v, const int* p, const double* m, double tau_) {
    N = N_; M = M_; K = K_; tau = S(tau_);
    view.assign(v, v + K); point.assign(p, p + K);
    meas.resize(2 * (size_t)K);
    for (size_t i = 0; i < 2 * (size_t)K; ++i) meas[i] = S(m[i]);
    pt_start.assign(M + 1, 0);
    for (int i = 0; i < K; ++i) pt_start[point[i] + 1]++;
    for (int j = 0; j < M; ++j) pt_start[j + 1] += pt_start[j];
    // block half-bandwidth of the reduced camera matrix from co-visibility
    bw = 0;
    for (int j = 0; j < M; ++j) {
      int lo = N, hi = -1;
      for (int i = pt_start[j]; i < pt_start[j + 1]; ++i) { lo = std::min(lo, view[i]); hi = std::max(hi, view[i]); }
      if (hi >= 0) bw = std::max(bw, hi - lo);
    }
    kd = 9 * bw + 8;
    if (kd > 9 * N - 1) kd = 9 * N - 1;
    R.resize(9 * (size_t)N); T.resize(3 * (size_t)N); f.resize(N); k1.resize(N); k2.resize(N);
    X.resize(3 * (size_t)M);
  }

  // ---------------------------------------------------------------- model (BAFunctor.h:151-178)
  static inline S psi(S tau2, S r2) { return (r2 < tau2) ? r2 * (S(2.0) - r2 / tau2) / S(4.0) : tau2 / S(4.0); }
  static inline S psi_weight(S tau2, S r2) { return std::max(S(0.0), S(1.0) - r2 / tau2); }

  struct Cam { const S* R; const S* T; S f, k1, k2; };
  Cam cam(const std::vector<S>& R_, const std::vector<S>& T_, const std::vector<S>& f_,
          const std::vector<S>& k1_, const std::vector<S>& k2_, int c) const {
    return Cam{&R_[9 * (size_t)c], &T_[3 * (size_t)c], f_[c], k1_[c], k2_[c]};
  }

  static inline void project(const Cam& c, const S* Xp, S q[2]) {
    // CameraMatrix.cpp:259-261, BAFunctor.h:151-156, DistortionFunction.cpp:14-23
    S XX[3];
    for (int r = 0; r < 3; ++r) XX[r] = c.R[3 * r] * Xp[0] + c.R[3 * r + 1] * Xp[1] + c.R[3 * r + 2] * Xp[2] + c.T[r];
    S xu0 = XX[0] / XX[2], xu1 = XX[1] / XX[2];
    S r2 = xu0 * xu0 + xu1 * xu1, r4 = r2 * r2;
    S kr = 1 + c.k1 * r2 + c.k2 * r4;
    q[0] = c.f * (kr * xu0);
    q[1] = c.f * (kr * xu1);
  }

  void residuals(const std::vector<S>& R_, const std::vector<S>& T_, const std::vector<S>& f_,
                 const std::vector<S>& k1_, const std::vector<S>& k2_, const std::vector<S>& X_,
                 std::vector<S>& fvec) const {
    // BAFunctor::E_pos, BAFunctor.h:160-178
    fvec.resize(2 * (size_t)K);
    const S tau2 = tau * tau;
#pragma omp parallel for if (nthreads > 1) num_threads(nthreads) schedule(static)
    for (int i = 0; i < K; ++i) {
      S q[2];
      project(cam(R_, T_, f_, k1_, k2_, view[i]), &X_[3 * (size_t)point[i]], q);
      S r0 = q[0] - meas[2 * (size_t)i], r1 = q[1] - meas[2 * (size_t)i + 1];
      S r2 = r0 * r0 + r1 * r1;
      S sqrt_psi = std::sqrt(psi(tau2, r2));
      S rnorm_r = S(1.0) / std::max(eps_psi, std::sqrt(r2));
      fvec[2 * (size_t)i] = r0 * sqrt_psi * rnorm_r;
      fvec[2 * (size_t)i + 1] = r1 * sqrt_psi * rnorm_r;
    }
  }

  static S sqnorm(const std::vector<S>& v) { S s = 0; for (S x : v) s += x * x; return s; }

  // ------------------------------------------------------------- Jacobian (BAFunctor.h:181-297)
  void jacobian_obs(int i, S* jc /*2x9*/, S* jp /*2x3*/) const {
    const Cam c = cam(R, T, f, k1, k2, view[i]);
    const S* Xp = &X[3 * (size_t)point[i]];
    S XX[3];
    for (int r = 0; r < 3; ++r) XX[r] = c.R[3 * r] * Xp[0] + c.R[3 * r + 1] * Xp[1] + c.R[3 * r + 2] * Xp[2] + c.T[r];
    // poseDerivatives (:126-142): d/dT = I ; d/domega = -[XX - T]_x ; d/dX = R
    S D[3] = {XX[0] - c.T[0], XX[1] - c.T[1], XX[2] - c.T[2]};
    S dRT[3][6] = {{1, 0, 0, 0, D[2], -D[1]}, {0, 1, 0, -D[2], 0, D[0]}, {0, 0, 1, D[1], -D[0], 0}};
    S xu[2] = {XX[0] / XX[2], XX[1] / XX[2]};
    S r2u = xu[0] * xu[0] + xu[1] * xu[1], r4u = r2u * r2u;
    S kr = 1 + c.k1 * r2u + c.k2 * r4u;
    S xd[2] = {kr * xu[0], kr * xu[1]};
    // dxu_dXX (:219-221)
    S dxu[2][3] = {{S(1.0) / XX[2], 0, -XX[0] / (XX[2] * XX[2])}, {0, S(1.0) / XX[2], -XX[1] / (XX[2] * XX[2])}};
    // dxd_dxu (DistortionFunction.cpp:38-51)
    S dkr = 2 * c.k1 + 4 * c.k2 * r2u;
    S dd[2][2];
    dd[0][0] = kr + xu[0] * xu[0] * dkr; dd[0][1] = xu[0] * xu[1] * dkr;
    dd[1][0] = dd[0][1];                 dd[1][1] = kr + xu[1] * xu[1] * dkr;
    S dpxu[2][2];  // dp_dxd * dxd_dxu with dp_dxd = f I
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) dpxu[a][b] = c.f * dd[a][b];
    S dpXX[2][3];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) dpXX[a][b] = dpxu[a][0] * dxu[0][b] + dpxu[a][1] * dxu[1][b];
    // outer derivative of the robust kernel (:227-242)
    const S tau2 = tau * tau;
    S q[2] = {c.f * xd[0], c.f * xd[1]};
    S rr[2] = {q[0] - meas[2 * (size_t)i], q[1] - meas[2 * (size_t)i + 1]};
    S r2 = rr[0] * rr[0] + rr[1] * rr[1];
    S W = psi_weight(tau2, r2);
    S sqrt_psi = std::sqrt(psi(tau2, r2));
    S rsqrt_psi = S(1.0) / std::max(eps_psi, sqrt_psi);
    S rcp_r2 = S(1.0) / std::max(eps_psi, r2);
    S rnorm_r = S(1.0) / std::max(eps_psi, std::sqrt(r2));
    S r_rt[2][2], rI[2][2] = {{std::sqrt(r2), 0}, {0, std::sqrt(r2)}}, outer[2][2];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) r_rt[a][b] = rr[a] * rr[b] * rnorm_r;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b)
      outer[a][b] = W / S(2.0) * rsqrt_psi * r_rt[a][b] + sqrt_psi * rcp_r2 * (rI[a][b] - r_rt[a][b]);
    // 2x12 block: cols 0-5 = dp_dXX*dXX_dRT, 6 = xd, 7-8 = f*dxd_dk, 9-11 = dp_dXX*R (:244-258)
    S blk[2][12];
    for (int a = 0; a < 2; ++a) {
      for (int b = 0; b < 6; ++b) blk[a][b] = dpXX[a][0] * dRT[0][b] + dpXX[a][1] * dRT[1][b] + dpXX[a][2] * dRT[2][b];
      blk[a][6] = xd[a];
      blk[a][7] = c.f * (xu[a] * r2u);
      blk[a][8] = c.f * (xu[a] * r4u);
      for (int b = 0; b < 3; ++b) blk[a][9 + b] = dpXX[a][0] * c.R[b] + dpXX[a][1] * c.R[3 + b] + dpXX[a][2] * c.R[6 + b];
    }
    for (int a = 0; a < 2; ++a) {
      for (int b = 0; b < 9; ++b) jc[9 * a + b] = outer[a][0] * blk[0][b] + outer[a][1] * blk[1][b];
      for (int b = 0; b < 3; ++b) jp[3 * a + b] = outer[a][0] * blk[0][9 + b] + outer[a][1] * blk[1][9 + b];
    }
  }

  // r = f(x), J = df(x), JtRes, column norms (QRChol.h:257-280; More.h:262-285; Cholesky.h:244-265)
  void linearize(double* energy, double* max_colnorm2, double* max_colnorm) {
    residuals(R, T, f, k1, k2, X, res);
    Jc.resize(18 * (size_t)K); Jp.resize(6 * (size_t)K);
#pragma omp parallel for if (nthreads > 1) num_threads(nthreads) schedule(static)
    for (int i = 0; i < K; ++i) jacobian_obs(i, &Jc[18 * (size_t)i], &Jp[6 * (size_t)i]);
    const int n = nparams();
    JtRes.assign(n, S(0));
    std::vector<S> cn2(n, S(0));
    for (int i = 0; i < K; ++i) {
      const S* jc = &Jc[18 * (size_t)i]; const S* jp = &Jp[6 * (size_t)i];
      const S e0 = res[2 * (size_t)i], e1 = res[2 * (size_t)i + 1];
      const size_t pc = 3 * (size_t)point[i], cc = 3 * (size_t)M + 9 * (size_t)view[i];
      for (int b = 0; b < 3; ++b) { JtRes[pc + b] -= jp[b] * e0 + jp[3 + b] * e1; cn2[pc + b] += jp[b] * jp[b] + jp[3 + b] * jp[3 + b]; }
      for (int b = 0; b < 9; ++b) { JtRes[cc + b] -= jc[b] * e0 + jc[9 + b] * e1; cn2[cc + b] += jc[b] * jc[b] + jc[9 + b] * jc[9 + b]; }
    }
    S mx = 0;
    for (int c = 0; c < n; ++c) mx = std::max(mx, cn2[c]);
    if (energy) *energy = (double)sqnorm(res);
    if (max_colnorm2) *max_colnorm2 = (double)mx;
    if (max_colnorm) *max_colnorm = (double)std::sqrt(mx);
  }

  // ------------------------------------------------------------ update (BAFunctor.h:299-342)
  static void rodrigues(const S w[3], S Rm[9]) {  // MathUtils.h:66-82 (hard 1e-6 cut-off, quirk Q2)
    S theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int i = 0; i < 9; ++i) Rm[i] = (i % 4 == 0) ? S(1) : S(0);
    if (std::abs(theta) > S(1e-6)) {
      S J[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}, J2[9];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) J2[3 * a + b] = J[3 * a] * J[b] + J[3 * a + 1] * J[3 + b] + J[3 * a + 2] * J[6 + b];
      const S c1 = std::sin(theta) / theta;
      const S c2 = (S(1.0) - std::cos(theta)) / (theta * theta);
      for (int i = 0; i < 9; ++i) Rm[i] = Rm[i] + c1 * J[i] + c2 * J2[i];
    }
  }

  void updated(const std::vector<S>& dx, std::vector<S>& R_, std::vector<S>& T_, std::vector<S>& f_,
               std::vector<S>& k1_, std::vector<S>& k2_, std::vector<S>& X_) const {
    R_ = R; T_ = T; f_ = f; k1_ = k1; k2_ = k2; X_ = X;
    const size_t cb = 3 * (size_t)M;
    for (int c = 0; c < N; ++c) {
      const S* p = &dx[cb + 9 * (size_t)c];
      for (int r = 0; r < 3; ++r) T_[3 * (size_t)c + r] += p[r];
      S dR[9], R0[9];
      rodrigues(p + 3, dR);
      std::memcpy(R0, &R_[9 * (size_t)c], sizeof(R0));
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
        R_[9 * (size_t)c + 3 * a + b] = dR[3 * a] * R0[b] + dR[3 * a + 1] * R0[3 + b] + dR[3 * a + 2] * R0[6 + b];
      k1_[c] += p[7]; k2_[c] += p[8]; f_[c] += p[6];
    }
    for (size_t i = 0; i < 3 * (size_t)M; ++i) X_[i] += dx[i];
  }

  double energy_at(const std::vector<S>& dx) const {
    std::vector<S> R_, T_, f_, k1_, k2_, X_, fv;
    updated(dx, R_, T_, f_, k1_, k2_, X_);
    residuals(R_, T_, f_, k1_, k2_, X_, fv);
    return (double)sqnorm(fv);
  }
  void apply(const std::vector<S>& dx) {
    std::vector<S> R_, T_, f_, k1_, k2_, X_;
    updated(dx, R_, T_, f_, k1_, k2_, X_);
    R.swap(R_); T.swap(T_); f.swap(f_); k1.swap(k1_); k2.swap(k2_); X.swap(X_);
  }

  // ------------------------------------------------ dense column-pivoted Householder QR (small)
  // Eigen ColPivHouseholderQR conventions: pivot = column of largest remaining norm; reflector
  // H = I - tau v v^T, v = [1; essential], beta = -sign(c0)*||x|| (makeHouseholder).
  // A: rows x 3 (row-major, ld 3), overwritten: R in the upper triangle, essentials below.
  struct QR3 { S tau[3]; int perm[3]; int rank; };
  static void householder_qr3(S* A, int rows, QR3& qr, bool pivot) {
    qr.perm[0] = 0; qr.perm[1] = 1; qr.perm[2] = 2; qr.rank = 0;
    S maxnorm2 = 0;
    const int steps = std::min(rows, 3);
    for (int k = 0; k < 3; ++k) qr.tau[k] = 0;
    for (int k = 0; k < steps; ++k) {
      if (pivot) {
        int best = k; S bn = -1;
        for (int c = k; c < 3; ++c) { S s = 0; for (int r = k; r < rows; ++r) s += A[3 * r + c] * A[3 * r + c]; if (s > bn) { bn = s; best = c; } }
        if (k == 0) maxnorm2 = bn;
        if (best != k) { for (int r = 0; r < rows; ++r) std::swap(A[3 * r + k], A[3 * r + best]); std::swap(qr.perm[k], qr.perm[best]); }
        if (bn > maxnorm2 * std::numeric_limits<S>::epsilon() * std::numeric_limits<S>::epsilon() * S(rows) && bn > 0) qr.rank = k + 1;
      } else qr.rank = k + 1;
      S c0 = A[3 * k + k], tail2 = 0;
      for (int r = k + 1; r < rows; ++r) tail2 += A[3 * r + k] * A[3 * r + k];
      S beta, tau_;
      if (tail2 <= std::numeric_limits<S>::min()) { tau_ = 0; beta = c0; for (int r = k + 1; r < rows; ++r) A[3 * r + k] = 0; }
      else {
        beta = std::sqrt(c0 * c0 + tail2);
        if (c0 >= 0) beta = -beta;
        const S inv = S(1) / (c0 - beta);
        for (int r = k + 1; r < rows; ++r) A[3 * r + k] *= inv;
        tau_ = (beta - c0) / beta;
      }
      A[3 * k + k] = beta; qr.tau[k] = tau_;
      for (int c = k + 1; c < 3; ++c) {  // apply H to the remaining columns
        S s = A[3 * k + c];
        for (int r = k + 1; r < rows; ++r) s += A[3 * r + k] * A[3 * r + c];
        s *= tau_;
        A[3 * k + c] -= s;
        for (int r = k + 1; r < rows; ++r) A[3 * r + c] -= s * A[3 * r + k];
      }
    }
  }
  // apply Q^T (H2 H1 H0 ...) to a dense block B (rows x ncols, row-major ld)
  static void apply_qt3(const S* A, int rows, const QR3& qr, S* B, int ncols, int ld) {
    const int steps = std::min(rows, 3);
    for (int k = 0; k < steps; ++k) {
      if (qr.tau[k] == S(0)) continue;
      for (int c = 0; c < ncols; ++c) {
        S s = B[(size_t)k * ld + c];
        for (int r = k + 1; r < rows; ++r) s += A[3 * r + k] * B[(size_t)r * ld + c];
        s *= qr.tau[k];
        B[(size_t)k * ld + c] -= s;
        for (int r = k + 1; r < rows; ++r) B[(size_t)r * ld + c] -= s * A[3 * r + k];
      }
    }
  }

  // -------------------------------------------------------------- reduced-system storage
  inline S& Sat(int i, int j) { return Sb[(size_t)i * (kd + 1) + (j - i + kd)]; }  // j <= i, i-j <= kd
  void clear_reduced() { Sb.assign((size_t)9 * N * (kd + 1), S(0)); g.assign(9 * (size_t)N, S(0)); }
  void keep_reduced() { S_last = Sb; g_last = g; }

  // Accumulation target of the per-point contributions: the reduced system itself (row0 = 0) or, in the
  // OpenMP timing leg, a chunk-local copy of the rows [row0, row0 + rows) that is added afterwards.
  struct Acc {
    S* Sb; S* g; int row0; int kd;
    inline S& at(int i, int j) { return Sb[(size_t)(i - row0) * (kd + 1) + (j - i + kd)]; }
    inline S& gat(int i) { return g[i - row0]; }
  };
  // Runs body(j, acc, A, B) for every point. nthreads > 1: contiguous chunks of points (balanced by n_j^3), each into
  // a chunk-local band copy covering the chunk's camera range, merged in chunk order afterwards.
  template <class Body>
  void for_points(Body&& body) {
    if (nthreads <= 1) {
      Acc acc{Sb.data(), g.data(), 0, kd};
      std::vector<S> A, B;
      for (int j = 0; j < M; ++j) body(j, acc, A, B);
      return;
    }
    const int nchunks = std::min(M, 8 * nthreads);
    std::vector<double> cum(M + 1, 0.0);
    for (int j = 0; j < M; ++j) { const double nj = pt_start[j + 1] - pt_start[j]; cum[j + 1] = cum[j] + nj * nj * nj + 20.0 * nj; }
    std::vector<int> cut(nchunks + 1, M);
    cut[0] = 0;
    for (int c = 1; c < nchunks; ++c) cut[c] = (int)(std::lower_bound(cum.begin(), cum.end(), cum[M] * c / nchunks) - cum.begin());
    for (int c = 1; c <= nchunks; ++c) cut[c] = std::max(cut[c], cut[c - 1]);
    std::vector<std::vector<S>> locS(nchunks), locg(nchunks);
    std::vector<int> row0(nchunks, 0), rows(nchunks, 0);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int c = 0; c < nchunks; ++c) {
      if (cut[c + 1] <= cut[c]) continue;
      int lo = N, hi = -1;
      for (int i = pt_start[cut[c]]; i < pt_start[cut[c + 1]]; ++i) { lo = std::min(lo, view[i]); hi = std::max(hi, view[i]); }
      row0[c] = 9 * lo; rows[c] = 9 * (hi - lo + 1);
      locS[c].assign((size_t)rows[c] * (kd + 1), S(0)); locg[c].assign(rows[c], S(0));
      Acc acc{locS[c].data(), locg[c].data(), row0[c], kd};
      std::vector<S> A, B;
      for (int j = cut[c]; j < cut[c + 1]; ++j) body(j, acc, A, B);
    }
    for (int c = 0; c < nchunks; ++c) {
      if (!rows[c]) continue;
      const S* ls = locS[c].data();
      S* dst = Sb.data() + (size_t)row0[c] * (kd + 1);
      const size_t cnt = (size_t)rows[c] * (kd + 1);
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (long long e = 0; e < (long long)cnt; ++e) dst[e] += ls[e];
      for (int r = 0; r < rows[c]; ++r) g[row0[c] + r] += locg[c][r];
    }
  }

  // un-pivoted LDL^T on the lower band + solve (SimplicialLDLT arithmetic in natural order;
  // D may be negative, QRChol.h:339-341)
  // Right-looking variant for the OpenMP timing leg: column j is scaled, then the rows below receive their rank-1
  // update in parallel (same arithmetic as band_ldlt_solve up to the summation order inside an entry).
  bool band_ldlt_solve_mt(std::vector<S>& y) {
    const int n = 9 * N;
    std::vector<S> d(n), col(kd + 1);
    for (int j = 0; j < n; ++j) {
      const S dj = Sat(j, j);
      d[j] = dj;
      if (dj == S(0) || !(dj == dj)) return false;
      const int i1 = std::min(n - 1, j + kd);
      for (int i = j + 1; i <= i1; ++i) col[i - j] = Sat(i, j);            // L_ij d_j
#pragma omp parallel for num_threads(nthreads) schedule(static) if (i1 - j > 64)
      for (int i = j + 1; i <= i1; ++i) {
        const S lij = col[i - j] / dj;
        S* row = &Sb[(size_t)i * (kd + 1) + kd - i];                       // row[k] = S(i, k)
        for (int k = j + 1; k <= i; ++k) row[k] -= lij * col[k - j];
        row[j] = lij;
      }
    }
    y = g;
    for (int i = 0; i < n; ++i) { S s = y[i]; for (int j = std::max(0, i - kd); j < i; ++j) s -= Sat(i, j) * y[j]; y[i] = s; }
    for (int i = 0; i < n; ++i) y[i] /= d[i];
    for (int i = n - 1; i >= 0; --i) {
      const S yi = y[i];
      for (int j = std::max(0, i - kd); j < i; ++j) y[j] -= Sat(i, j) * yi;
    }
    return true;
  }

  bool band_ldlt_solve(std::vector<S>& y) {
    if (nthreads > 1) return band_ldlt_solve_mt(y);
    const int n = 9 * N;
    std::vector<S> d(n);
    for (int i = 0; i < n; ++i) {
      const int j0 = std::max(0, i - kd);
      for (int j = j0; j < i; ++j) {
        S s = Sat(i, j);
        const int k0 = std::max(j0, j - kd);
        for (int k = k0; k < j; ++k) s -= Sat(i, k) * Sat(j, k);  // Sat(i,k) = L_ik*d_k (this row, temporary), Sat(j,k) = L_jk
        Sat(i, j) = s;  // temporarily L_ij * d_j
      }
      S di = Sat(i, i);
      for (int j = j0; j < i; ++j) { const S lij = Sat(i, j) / d[j]; di -= lij * Sat(i, j); Sat(i, j) = lij; }
      d[i] = di; Sat(i, i) = di;
      if (di == S(0) || !(di == di)) return false;
    }
    y = g;
    for (int i = 0; i < n; ++i) { S s = y[i]; for (int j = std::max(0, i - kd); j < i; ++j) s -= Sat(i, j) * y[j]; y[i] = s; }
    for (int i = 0; i < n; ++i) y[i] /= d[i];
    for (int i = n - 1; i >= 0; --i) {
      const S yi = y[i];
      for (int j = std::max(0, i - kd); j < i; ++j) y[j] -= Sat(i, j) * yi;
    }
    return true;
  }

  // Householder QR solve of the square symmetric band matrix S y = g (right block of QRKIT/MOREQR
  // restated on the reduced camera matrix; see DESIGN.md for the deviation from the tall J2bot QR)
  void band_qr_solve(std::vector<S>& y) {
    const int n = 9 * N;
    const int ku = std::min(n - 1, 2 * kd), ld = kd + ku + 1;
    std::vector<S> G((size_t)n * ld, S(0));
    auto Gat = [&](int i, int j) -> S& { return G[(size_t)i * ld + (j - i + kd)]; };
    for (int i = 0; i < n; ++i) for (int j = std::max(0, i - kd); j <= i; ++j) { Gat(i, j) = Sat(i, j); Gat(j, i) = Sat(i, j); }
    y = g;
    for (int k = 0; k < n; ++k) {
      const int r1 = std::min(n - 1, k + kd), c1 = std::min(n - 1, k + ku);
      S c0 = Gat(k, k), tail2 = 0;
      for (int r = k + 1; r <= r1; ++r) tail2 += Gat(r, k) * Gat(r, k);
      if (tail2 <= std::numeric_limits<S>::min()) continue;
      S beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0) beta = -beta;
      const S inv = S(1) / (c0 - beta), tau_ = (beta - c0) / beta;
      for (int r = k + 1; r <= r1; ++r) Gat(r, k) *= inv;
      Gat(k, k) = beta;
#pragma omp parallel for if (nthreads > 1 && c1 - k > 64) num_threads(nthreads) schedule(static)
      for (int c = k + 1; c <= c1; ++c) {
        S s = Gat(k, c);
        for (int r = k + 1; r <= r1; ++r) s += Gat(r, k) * Gat(r, c);
        s *= tau_;
        Gat(k, c) -= s;
        for (int r = k + 1; r <= r1; ++r) Gat(r, c) -= s * Gat(r, k);
      }
      S s = y[k];
      for (int r = k + 1; r <= r1; ++r) s += Gat(r, k) * y[r];
      s *= tau_;
      y[k] -= s;
      for (int r = k + 1; r <= r1; ++r) y[r] -= s * Gat(r, k);
    }
    for (int i = n - 1; i >= 0; --i) {
      S s = y[i];
      for (int j = i + 1; j <= std::min(n - 1, i + ku); ++j) s -= Gat(i, j) * y[j];
      y[i] = s / Gat(i, i);
    }
  }

  // ----------------------------------------------------------------------------------- steps
  // QRCHOL / QRKIT: QRChol.h:282-360. Per point j the block [Jp_j; sqrt(lambda) I3] (App. C) is
  // factored, Q_j^T applied to [Jc_j; 0] and [r_j; 0]; top 3 rows -> R12_j, c_j; bottom 2n_j rows
  // -> J2bot_j, d_j. S = sum J2bot_j^T J2bot_j + lambda I (camera lambda rows), g = sum J2bot_j^T d_j.
  // y = S^-1 g ; dx_cam = -y ; dx_j = -R_j^-1 (c_j - R12_j y).
  bool step_qr(Variant variant, S lambda, std::vector<S>& dx) {
    const int n = nparams();
    dx.assign(n, S(0));
    clear_reduced();
    const S sl = std::sqrt(lambda);
    std::vector<S> Rj(9 * (size_t)M), cj(3 * (size_t)M);
    std::vector<int> permj(3 * (size_t)M);
    std::vector<S> R12all(27 * (size_t)K);
    // optional reference-faithful tall QR of J2bot (QRKIT, small problems only; single-threaded)
    std::vector<S> tall; std::vector<S> tall_rhs; size_t tall_rows = 0; const int nc = 9 * N;
    if (variant == QRKIT && tall_qr) { tall.assign(((size_t)2 * K + nc) * nc, S(0)); tall_rhs.assign((size_t)2 * K + nc, S(0)); }
    const int nthreads_saved = nthreads;
    if (!tall.empty()) nthreads = 1;
    for_points([&](int j, Acc& acc, std::vector<S>& A, std::vector<S>& B) {
      const int o0 = pt_start[j], nj = pt_start[j + 1] - o0, rows = 2 * nj + 3, ncols = 9 * nj;
      A.assign((size_t)rows * 3, S(0));
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) A[3 * (2 * i + a) + b] = Jp[6 * (size_t)(o0 + i) + 3 * a + b];
      for (int b = 0; b < 3; ++b) A[3 * (2 * nj + b) + b] = sl;
      QR3 qr; householder_qr3(A.data(), rows, qr, true);
      // B = [blockdiag(Jc_i) | r ; 0]
      const int ld = ncols + 1;
      B.assign((size_t)rows * ld, S(0));
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 2; ++a) {
        for (int b = 0; b < 9; ++b) B[(size_t)(2 * i + a) * ld + 9 * i + b] = Jc[18 * (size_t)(o0 + i) + 9 * a + b];
        B[(size_t)(2 * i + a) * ld + ncols] = res[2 * (size_t)(o0 + i) + a];
      }
      apply_qt3(A.data(), rows, qr, B.data(), ld, ld);
      for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) Rj[9 * (size_t)j + 3 * a + b] = (b >= a) ? A[3 * a + b] : S(0); cj[3 * (size_t)j + a] = B[(size_t)a * ld + ncols]; permj[3 * (size_t)j + a] = qr.perm[a]; }
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 3; ++a) for (int b = 0; b < 9; ++b) R12all[27 * (size_t)(o0 + i) + 9 * a + b] = B[(size_t)a * ld + 9 * i + b];
      if (!tall.empty()) {
        for (int r = 3; r < rows; ++r) {
          for (int i = 0; i < nj; ++i) for (int b = 0; b < 9; ++b) tall[(tall_rows) * nc + 9 * view[o0 + i] + b] = B[(size_t)r * ld + 9 * i + b];
          tall_rhs[tall_rows] = B[(size_t)r * ld + ncols];
          ++tall_rows;
        }
      }
      // S += J2bot_j^T J2bot_j (lower), g += J2bot_j^T d_j   [explicit product, QRChol.h:339-341]
      for (int ia = 0; ia < nj; ++ia) for (int ib = 0; ib <= ia; ++ib) {
        int ca = view[o0 + ia], cb = view[o0 + ib], xa = ia, xb = ib;
        if (ca < cb) { std::swap(ca, cb); std::swap(xa, xb); }
        for (int p = 0; p < 9; ++p) for (int q = 0; q < 9; ++q) {
          if (ca == cb && q > p) continue;
          S s = 0;
          for (int r = 3; r < rows; ++r) s += B[(size_t)r * ld + 9 * xa + p] * B[(size_t)r * ld + 9 * xb + q];
          acc.at(9 * ca + p, 9 * cb + q) += s;
        }
      }
      for (int ia = 0; ia < nj; ++ia) for (int p = 0; p < 9; ++p) {
        S s = 0;
        for (int r = 3; r < rows; ++r) s += B[(size_t)r * ld + 9 * ia + p] * B[(size_t)r * ld + ncols];
        acc.gat(9 * view[o0 + ia] + p) += s;
      }
    });
    nthreads = nthreads_saved;
    for (int i = 0; i < 9 * N; ++i) Sat(i, i) += sl * sl;  // camera lambda rows at the bottom of J2bot
    keep_reduced();
    std::vector<S> y;
    if (variant == QRCHOL) { if (!band_ldlt_solve(y)) return false; }
    else if (!tall.empty()) {
      for (int i = 0; i < nc; ++i) { tall[(tall_rows) * nc + i] = sl; ++tall_rows; }
      dense_qr_solve_tall(tall, tall_rhs, (int)tall_rows, nc, y);
    } else band_qr_solve(y);
    backsubstitute(Rj, cj, permj, R12all, y, dx);
    return true;
  }

  // reference-faithful dense Householder QR least squares (tall), used only for the QRKIT deviation test
  static void dense_qr_solve_tall(std::vector<S>& Am, std::vector<S>& b, int rows, int cols, std::vector<S>& y) {
    for (int k = 0; k < cols; ++k) {
      S c0 = Am[(size_t)k * cols + k], tail2 = 0;
      for (int r = k + 1; r < rows; ++r) tail2 += Am[(size_t)r * cols + k] * Am[(size_t)r * cols + k];
      if (tail2 <= std::numeric_limits<S>::min()) continue;
      S beta = std::sqrt(c0 * c0 + tail2); if (c0 >= 0) beta = -beta;
      const S inv = S(1) / (c0 - beta), tau_ = (beta - c0) / beta;
      for (int r = k + 1; r < rows; ++r) Am[(size_t)r * cols + k] *= inv;
      Am[(size_t)k * cols + k] = beta;
      std::vector<S> w(cols - k - 1, S(0));
      for (int c = k + 1; c < cols; ++c) w[c - k - 1] = Am[(size_t)k * cols + c];
      for (int r = k + 1; r < rows; ++r) { const S v = Am[(size_t)r * cols + k]; if (v == S(0)) continue; const S* row = &Am[(size_t)r * cols]; for (int c = k + 1; c < cols; ++c) w[c - k - 1] += v * row[c]; }
      for (auto& x : w) x *= tau_;
      for (int c = k + 1; c < cols; ++c) Am[(size_t)k * cols + c] -= w[c - k - 1];
      for (int r = k + 1; r < rows; ++r) { const S v = Am[(size_t)r * cols + k]; if (v == S(0)) continue; S* row = &Am[(size_t)r * cols]; for (int c = k + 1; c < cols; ++c) row[c] -= w[c - k - 1] * v; }
      S s = b[k]; for (int r = k + 1; r < rows; ++r) s += Am[(size_t)r * cols + k] * b[r];
      s *= tau_; b[k] -= s; for (int r = k + 1; r < rows; ++r) b[r] -= s * Am[(size_t)r * cols + k];
    }
    y.assign(cols, S(0));
    for (int i = cols - 1; i >= 0; --i) { S s = b[i]; for (int j = i + 1; j < cols; ++j) s -= Am[(size_t)i * cols + j] * y[j]; y[i] = s / Am[(size_t)i * cols + i]; }
  }

  // QRChol.h:344-360: R dx = -qtb with the right block replaced by [R12; I]; column un-permutation
  void backsubstitute(const std::vector<S>& Rj, const std::vector<S>& cj, const std::vector<int>& permj,
                      const std::vector<S>& R12all, const std::vector<S>& y, std::vector<S>& dx) const {
    for (int c = 0; c < 9 * N; ++c) dx[3 * (size_t)M + c] = -y[c];
#pragma omp parallel for if (nthreads > 1) num_threads(nthreads) schedule(static)
    for (int j = 0; j < M; ++j) {
      S rhs[3] = {-cj[3 * (size_t)j], -cj[3 * (size_t)j + 1], -cj[3 * (size_t)j + 2]};
      for (int i = pt_start[j]; i < pt_start[j + 1]; ++i) {
        const S* r12 = &R12all[27 * (size_t)i]; const S* yc = &y[9 * (size_t)view[i]];
        for (int a = 0; a < 3; ++a) { S s = 0; for (int b = 0; b < 9; ++b) s += r12[9 * a + b] * yc[b]; rhs[a] += s; }
      }
      const S* Rm = &Rj[9 * (size_t)j];
      S z[3];
      z[2] = rhs[2] / Rm[8];
      z[1] = (rhs[1] - Rm[5] * z[2]) / Rm[4];
      z[0] = (rhs[0] - Rm[1] * z[1] - Rm[2] * z[2]) / Rm[0];
      for (int a = 0; a < 3; ++a) dx[3 * (size_t)j + permj[3 * (size_t)j + a]] = z[a];
    }
  }

  // CHOLESKY: (J^T J + lambda I) dx = -J^T r by LDL^T (Cholesky.h:260-285), points eliminated first
  // (the fill-reducing order a minimum-degree ordering picks for this arrow structure). Quirk Q1
  // (the extra permutationP() at :285) is NOT reproduced — see DESIGN.md.
  bool step_cholesky(S lambda, std::vector<S>& dx) {
    const int n = nparams();
    dx.assign(n, S(0));
    clear_reduced();
    std::vector<S> Lj(3 * (size_t)M), Dj(3 * (size_t)M), zall(3 * (size_t)M), U(27 * (size_t)K);
    for (int j = 0; j < M; ++j) {
      const int o0 = pt_start[j], nj = pt_start[j + 1] - o0;
      S V[6] = {lambda, 0, lambda, 0, 0, lambda};  // lower: 00,10,11,20,21,22
      for (int i = 0; i < nj; ++i) {
        const S* jp = &Jp[6 * (size_t)(o0 + i)];
        V[0] += jp[0] * jp[0] + jp[3] * jp[3]; V[1] += jp[1] * jp[0] + jp[4] * jp[3]; V[2] += jp[1] * jp[1] + jp[4] * jp[4];
        V[3] += jp[2] * jp[0] + jp[5] * jp[3]; V[4] += jp[2] * jp[1] + jp[5] * jp[4]; V[5] += jp[2] * jp[2] + jp[5] * jp[5];
      }
      const S d0 = V[0], l10 = V[1] / d0, l20 = V[3] / d0;
      const S d1 = V[2] - l10 * l10 * d0, l21 = (V[4] - l20 * l10 * d0) / d1;
      const S d2 = V[5] - l20 * l20 * d0 - l21 * l21 * d1;
      Lj[3 * (size_t)j] = l10; Lj[3 * (size_t)j + 1] = l20; Lj[3 * (size_t)j + 2] = l21;
      Dj[3 * (size_t)j] = d0; Dj[3 * (size_t)j + 1] = d1; Dj[3 * (size_t)j + 2] = d2;
      // z = L^-1 b_p  with b_p = JtRes_p
      S z[3] = {JtRes[3 * (size_t)j], JtRes[3 * (size_t)j + 1], JtRes[3 * (size_t)j + 2]};
      z[1] -= l10 * z[0]; z[2] -= l20 * z[0] + l21 * z[1];
      for (int a = 0; a < 3; ++a) zall[3 * (size_t)j + a] = z[a];
      // U_i = L^-1 W_i^T (3x9), W_i = Jc_i^T Jp_i
      for (int i = 0; i < nj; ++i) {
        const S* jc = &Jc[18 * (size_t)(o0 + i)]; const S* jp = &Jp[6 * (size_t)(o0 + i)];
        S* u = &U[27 * (size_t)(o0 + i)];
        for (int b = 0; b < 9; ++b) {
          S w0 = jp[0] * jc[b] + jp[3] * jc[9 + b], w1 = jp[1] * jc[b] + jp[4] * jc[9 + b], w2 = jp[2] * jc[b] + jp[5] * jc[9 + b];
          w1 -= l10 * w0; w2 -= l20 * w0 + l21 * w1;
          u[b] = w0; u[9 + b] = w1; u[18 + b] = w2;
        }
      }
      const S id[3] = {S(1) / d0, S(1) / d1, S(1) / d2};
      for (int ia = 0; ia < nj; ++ia) for (int ib = 0; ib <= ia; ++ib) {
        int ca = view[o0 + ia], cb = view[o0 + ib], xa = ia, xb = ib;
        if (ca < cb) { std::swap(ca, cb); std::swap(xa, xb); }
        const S* ua = &U[27 * (size_t)(o0 + xa)]; const S* ub = &U[27 * (size_t)(o0 + xb)];
        const S* jca = &Jc[18 * (size_t)(o0 + xa)];
        for (int p = 0; p < 9; ++p) for (int q = 0; q < 9; ++q) {
          if (ca == cb && q > p) continue;
          S s = -(ua[p] * id[0] * ub[q] + ua[9 + p] * id[1] * ub[9 + q] + ua[18 + p] * id[2] * ub[18 + q]);
          if (xa == xb) s += jca[p] * jca[q] + jca[9 + p] * jca[9 + q];
          Sat(9 * ca + p, 9 * cb + q) += s;
        }
      }
      for (int i = 0; i < nj; ++i) {
        const S* u = &U[27 * (size_t)(o0 + i)];
        for (int p = 0; p < 9; ++p) g[9 * (size_t)view[o0 + i] + p] -= u[p] * id[0] * z[0] + u[9 + p] * id[1] * z[1] + u[18 + p] * id[2] * z[2];
      }
    }
    for (int c = 0; c < 9 * N; ++c) { Sat(c, c) += lambda; g[c] += JtRes[3 * (size_t)M + c]; }
    keep_reduced();
    std::vector<S> y;
    if (!band_ldlt_solve(y)) return false;  // y = dx_cam
    for (int c = 0; c < 9 * N; ++c) dx[3 * (size_t)M + c] = y[c];
    for (int j = 0; j < M; ++j) {
      S z[3] = {zall[3 * (size_t)j], zall[3 * (size_t)j + 1], zall[3 * (size_t)j + 2]};
      for (int i = pt_start[j]; i < pt_start[j + 1]; ++i) {
        const S* u = &U[27 * (size_t)i]; const S* yc = &y[9 * (size_t)view[i]];
        for (int a = 0; a < 3; ++a) { S s = 0; for (int b = 0; b < 9; ++b) s += u[9 * a + b] * yc[b]; z[a] -= s; }
      }
      for (int a = 0; a < 3; ++a) z[a] /= Dj[3 * (size_t)j + a];
      const S l10 = Lj[3 * (size_t)j], l20 = Lj[3 * (size_t)j + 1], l21 = Lj[3 * (size_t)j + 2];
      z[1] -= l21 * z[2]; z[0] -= l10 * z[1] + l20 * z[2];
      for (int a = 0; a < 3; ++a) dx[3 * (size_t)j + a] = z[a];
    }
    return true;
  }

  // MOREQR stage 1 (More.h:288-291): QR of the UN-damped J once per outer iteration.
  void moreqr_outer() {
    m_R.assign(9 * (size_t)M, S(0)); m_c.assign(3 * (size_t)M, S(0)); m_perm.assign(3 * (size_t)M, 0);
    m_R12.assign(27 * (size_t)K, S(0));
    clear_reduced();
    std::vector<S> A, B;
    for (int j = 0; j < M; ++j) {
      const int o0 = pt_start[j], nj = pt_start[j + 1] - o0, rows = std::max(2 * nj, 3), ncols = 9 * nj, ld = ncols + 1;
      A.assign((size_t)rows * 3, S(0));
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) A[3 * (2 * i + a) + b] = Jp[6 * (size_t)(o0 + i) + 3 * a + b];
      QR3 qr; householder_qr3(A.data(), rows, qr, true);
      B.assign((size_t)rows * ld, S(0));
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 2; ++a) {
        for (int b = 0; b < 9; ++b) B[(size_t)(2 * i + a) * ld + 9 * i + b] = Jc[18 * (size_t)(o0 + i) + 9 * a + b];
        B[(size_t)(2 * i + a) * ld + ncols] = res[2 * (size_t)(o0 + i) + a];
      }
      apply_qt3(A.data(), rows, qr, B.data(), ld, ld);
      for (int a = 0; a < 3; ++a) { for (int b = a; b < 3; ++b) m_R[9 * (size_t)j + 3 * a + b] = A[3 * a + b]; m_c[3 * (size_t)j + a] = B[(size_t)a * ld + ncols]; m_perm[3 * (size_t)j + a] = qr.perm[a]; }
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 3; ++a) for (int b = 0; b < 9; ++b) m_R12[27 * (size_t)(o0 + i) + 9 * a + b] = B[(size_t)a * ld + 9 * i + b];
      for (int ia = 0; ia < nj; ++ia) for (int ib = 0; ib <= ia; ++ib) {
        int ca = view[o0 + ia], cb = view[o0 + ib], xa = ia, xb = ib;
        if (ca < cb) { std::swap(ca, cb); std::swap(xa, xb); }
        for (int p = 0; p < 9; ++p) for (int q = 0; q < 9; ++q) {
          if (ca == cb && q > p) continue;
          S s = 0;
          for (int r = 3; r < rows; ++r) s += B[(size_t)r * ld + 9 * xa + p] * B[(size_t)r * ld + 9 * xb + q];
          Sat(9 * ca + p, 9 * cb + q) += s;
        }
      }
      for (int ia = 0; ia < nj; ++ia) for (int p = 0; p < 9; ++p) {
        S s = 0;
        for (int r = 3; r < rows; ++r) s += B[(size_t)r * ld + 9 * ia + p] * B[(size_t)r * ld + ncols];
        g[9 * (size_t)view[o0 + ia] + p] += s;
      }
    }
    m_S0 = Sb; m_g0 = g;
  }

  // MOREQR stage 2 (More.h:293-348): per lambda, QR of [R; sqrt(lambda) I]; per point a 6x3 block.
  bool step_moreqr(S lambda, std::vector<S>& dx) {
    const int n = nparams();
    dx.assign(n, S(0));
    Sb = m_S0; g = m_g0;
    const S sl = std::sqrt(lambda);
    std::vector<S> Rj(9 * (size_t)M), cj(3 * (size_t)M), R12n(27 * (size_t)K);
    std::vector<int> permj(3 * (size_t)M);
    std::vector<S> B;
    for (int j = 0; j < M; ++j) {
      const int o0 = pt_start[j], nj = pt_start[j + 1] - o0, ncols = 9 * nj, ld = ncols + 1;
      S A[18];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { A[3 * a + b] = m_R[9 * (size_t)j + 3 * a + b]; A[3 * (3 + a) + b] = (a == b) ? sl : S(0); }
      QR3 qr; householder_qr3(A, 6, qr, true);
      B.assign((size_t)6 * ld, S(0));
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 3; ++a) for (int b = 0; b < 9; ++b) B[(size_t)a * ld + 9 * i + b] = m_R12[27 * (size_t)(o0 + i) + 9 * a + b];
      for (int a = 0; a < 3; ++a) B[(size_t)a * ld + ncols] = m_c[3 * (size_t)j + a];
      apply_qt3(A, 6, qr, B.data(), ld, ld);
      for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) Rj[9 * (size_t)j + 3 * a + b] = (b >= a) ? A[3 * a + b] : S(0); cj[3 * (size_t)j + a] = B[(size_t)a * ld + ncols]; permj[3 * (size_t)j + a] = qr.perm[a]; }
      for (int i = 0; i < nj; ++i) for (int a = 0; a < 3; ++a) for (int b = 0; b < 9; ++b) R12n[27 * (size_t)(o0 + i) + 9 * a + b] = B[(size_t)a * ld + 9 * i + b];
      // fill rows F_j = rows 3..5 join the camera block: S += F^T F, g += F^T phi
      for (int ia = 0; ia < nj; ++ia) for (int ib = 0; ib <= ia; ++ib) {
        int ca = view[o0 + ia], cb = view[o0 + ib], xa = ia, xb = ib;
        if (ca < cb) { std::swap(ca, cb); std::swap(xa, xb); }
        for (int p = 0; p < 9; ++p) for (int q = 0; q < 9; ++q) {
          if (ca == cb && q > p) continue;
          S s = 0;
          for (int r = 3; r < 6; ++r) s += B[(size_t)r * ld + 9 * xa + p] * B[(size_t)r * ld + 9 * xb + q];
          Sat(9 * ca + p, 9 * cb + q) += s;
        }
      }
      for (int ia = 0; ia < nj; ++ia) for (int p = 0; p < 9; ++p) {
        S s = 0;
        for (int r = 3; r < 6; ++r) s += B[(size_t)r * ld + 9 * ia + p] * B[(size_t)r * ld + ncols];
        g[9 * (size_t)view[o0 + ia] + p] += s;
      }
    }
    for (int i = 0; i < 9 * N; ++i) Sat(i, i) += sl * sl;
    keep_reduced();
    std::vector<S> y;
    band_qr_solve(y);
    // back-substitute in the doubly-permuted point ordering (More.h:344-348)
    std::vector<S> dxs(n, S(0));
    std::vector<int> ident(3 * (size_t)M);
    for (size_t i = 0; i < ident.size(); ++i) ident[i] = (int)(i % 3);
    // inner permutation first (solution is in inner-pivot order), then outer
    backsubstitute(Rj, cj, permj, R12n, y, dxs);
    for (int c = 0; c < 9 * N; ++c) dx[3 * (size_t)M + c] = dxs[3 * (size_t)M + c];
    for (int j = 0; j < M; ++j) for (int a = 0; a < 3; ++a) dx[3 * (size_t)j + m_perm[3 * (size_t)j + a]] = dxs[3 * (size_t)j + a];
    return true;
  }

  bool step(Variant v, double lambda, std::vector<S>& dx) {
    switch (v) {
      case QRKIT: case QRCHOL: return step_qr(v, S(lambda), dx);
      case CHOLESKY: return step_cholesky(S(lambda), dx);
      case MOREQR: return step_moreqr(S(lambda), dx);
    }
    return false;
  }

  // --------------------------------------------------------------- LM loops (one per variant)
  // QRChol.h:204-436 / More.h:204-425 / Cholesky.h:190-361. QRKIT's own loop is NOT IN TREE:
  // assumption = QRCHOL control flow and lambda_0 rule.
  Status minimize(Variant variant, int max_outer, std::vector<TrialRecord>& log) {
    const S lam_min = S(1e-10), lam_max = S(1e10), inc_base = S(2);  // QRChol.h:131-133
    const S tolFun = S(1e-8);                                        // :143
    const int maxIter = (int)1e6, maxFunEv = (int)1e6;               // :144-145
    S lambda = S(1e-3), lambdaInc = inc_base;
    int funEvals = 0, iter = 0;
    S hist[2] = {0, 0};
    Status status = Running;
    std::vector<S> dx;
    bool stopNow = false;
    while (true) {
      iter++;
      if (iter > maxIter || (max_outer > 0 && iter > max_outer)) { status = MaxItersReached; break; }
      if (funEvals > maxFunEv) { status = TooManyFunctionEvaluation; break; }
      double e0, cn2, cn;
      linearize(&e0, &cn2, &cn);
      funEvals++;
      S energy = S(e0);
      if (iter == 1) lambda = (variant == MOREQR) ? S(1e-6 * (double)S(cn)) : S(1e-12 * (double)S(cn2));  // quirk Q9
      if (variant == MOREQR) moreqr_outer();
      S energyTest = 0;
      while (true) {
        TrialRecord rec{}; rec.iter = iter; rec.energy = (double)energy; rec.lambda_used = (double)lambda;
        bool ok = step(variant, (double)lambda, dx);
        S dxn = 0; for (S v : dx) dxn += v * v;
        rec.dx_norm = std::sqrt((double)dxn);
        energyTest = ok ? S(energy_at(dx)) : std::numeric_limits<S>::quiet_NaN();
        funEvals++;
        rec.energy_test = (double)energyTest;
        if (energyTest < energy) {
          S rhoScale = 0;
          for (size_t i = 0; i < dx.size(); ++i) rhoScale += dx[i] * (lambda * dx[i] + JtRes[i]);
          S rho = (energy - energyTest) / rhoScale;
          S lambdaMul = S(1.0) - std::pow(S(2.0) * rho - S(1.0), S(3.0));
          lambda *= std::max<S>(S(1.0) / S(3.0), lambdaMul);
          lambda = std::max<S>(lambda, lam_min);
          rec.accepted = 1; rec.rho = (double)rho; rec.lambda_next = (double)lambda;
          log.push_back(rec);
          lambdaInc = inc_base;
          energy = energyTest;
          hist[iter % 2] = energy;
          break;
        } else {
          rec.accepted = 0; rec.rho = 0; rec.lambda_next = (double)lambda;
          log.push_back(rec);
          if (lambda > lam_max) { status = ExceededLambdaMax; stopNow = true; break; }
          lambda *= lambdaInc;
          lambdaInc = (variant == QRCHOL || variant == QRKIT) ? S(std::pow(lambdaInc, 1.5f)) : S(std::pow(lambdaInc, 1.5));  // quirk Q5
        }
      }
      if (stopNow) break;
      if (iter > 2) {
        S maxf = std::max(hist[0], hist[1]);
        if (std::abs(energy - maxf) < tolFun * energy) { status = Success; break; }  // x NOT committed (Q8)
      }
      apply(dx);
    }
    return status;
  }
};

}  // namespace bao

// GPU-backed mirror of the reference functor (src/Optimization/BAFunctor.{h,cpp}): same members the
// LM loops use — InputType, operator() (energy), df (linearisation), increment/accept, and the
// QRSolver typedef selected by the QRKIT / QRCHOL / MOREQR / CHOLESKY symbols (BAFunctor.h:98-119) —
// with all arithmetic behind the C ABI of include/ba_gpu.h. No CPU fallback.
#ifndef BA_FUNCTOR_H
#define BA_FUNCTOR_H
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ba_gpu.h"
#include "BATypeUtils.h"

#if !defined(QRKIT) && !defined(QRCHOL) && !defined(MOREQR) && !defined(CHOLESKY)
#error "define one of QRKIT, QRCHOL, MOREQR, CHOLESKY (src/CMakeLists.txt:109-160)"
#endif

inline void ba_check(int rc) { if (rc != 0) throw std::runtime_error(std::string("ba_gpu: ") + ba_last_error()); }

// The solver concept of the reference's LM loops, on the GPU (BAFunctor.h:98-119 typedef block).
struct GpuSchurSolver {
  ba_handle* h = nullptr;
  long blockRows = 0, blockCols = 0;
  void setSparseBlockParams(long r, long c) { blockRows = r; blockCols = c; }  // BAFunctor.cpp:66-76
  void compute(Scalar lambda) { ba_check(ba_compute(h, (double)lambda)); }     // QRChol.h:291-339
  void solve(Scalar& dxNorm, Scalar& rhoScale, Scalar& energyTest) {           // QRChol.h:322-371
    double a, b, c;
    ba_check(ba_solve_try(h, &a, &b, &c));
    dxNorm = (Scalar)a; rhoScale = (Scalar)b; energyTest = (Scalar)c;
  }
};

struct BAFunctor {
  // Variables for optimization live in InputType (BAFunctor.h:39-51): R instead of CameraMatrix objects
  struct InputType {
    std::vector<double> R, T, f, k1, k2, X;  // 9N, 3N, N, N, N, 3M (doubles cross the C ABI)
    long nCameras() const { return (long)f.size(); }
    long nDataPoints() const { return (long)X.size() / 3; }
  };
  typedef GpuSchurSolver SchurlikeQRSolver;
  typedef SchurlikeQRSolver QRSolver;

#if defined(QRKIT)
  static constexpr int kVariant = BA_QRKIT;
#elif defined(QRCHOL)
  static constexpr int kVariant = BA_QRCHOL;
#elif defined(MOREQR)
  static constexpr int kVariant = BA_MOREQR;
#else
  static constexpr int kVariant = BA_CHOLESKY;
#endif

  ba_handle* gpu = nullptr;
  const long numPoints, numCameras, numMeasurements;
  const long numParameters, numResiduals, numPointParams;

  BAFunctor(long numPoints_, long numCameras_, const std::vector<double>& measurements /*2K*/,
            const std::vector<int>& correspondingView, const std::vector<int>& correspondingPoint,
            Scalar inlierThreshold, int device = 0)
      : numPoints(numPoints_), numCameras(numCameras_), numMeasurements((long)correspondingView.size()),
        numParameters(numPoints_ * 3 + numCameras_ * 9), numResiduals((long)correspondingView.size() * 2),
        numPointParams(numPoints_ * 3) {
    ba_check(ba_create(&gpu, (int)numCameras, (int)numPoints, (int)numMeasurements, correspondingView.data(),
                       correspondingPoint.data(), measurements.data(), (double)inlierThreshold,
                       sizeof(Scalar) == 4 ? BA_F32 : BA_F64, kVariant, device));
  }
  BAFunctor(const BAFunctor&) = delete;
  ~BAFunctor() { if (gpu) ba_destroy(gpu); }

  long inputs() const { return numParameters; }
  long values() const { return numResiduals; }

  void upload(const InputType& x) { ba_check(ba_set_state(gpu, x.R.data(), x.T.data(), x.f.data(), x.k1.data(), x.k2.data(), x.X.data())); }
  void download(InputType& x) { ba_check(ba_get_state(gpu, x.R.data(), x.T.data(), x.f.data(), x.k1.data(), x.k2.data(), x.X.data())); }

  // 1. residual energy at the device-resident x   (operator(), BAFunctor.cpp:82-86)
  int operator()(Scalar& energy) { double e; ba_check(ba_eval(gpu, &e)); energy = (Scalar)e; return 0; }
  // 2. linearise at x: energy, max column norms    (df, BAFunctor.cpp:89-101 + QRChol.h:267-280)
  int df(Scalar& energy, Scalar* maxColNorm2, Scalar* maxColNorm) {
    double e, a, b;
    ba_check(ba_linearize(gpu, &e, maxColNorm2 ? &a : nullptr, maxColNorm ? &b : nullptr));
    energy = (Scalar)e;
    if (maxColNorm2) *maxColNorm2 = (Scalar)a;
    if (maxColNorm) *maxColNorm = (Scalar)b;
    return 0;
  }
  void accept() { ba_check(ba_accept(gpu)); }   // x = xTest, QRChol.h:428
  void reject() { ba_check(ba_reject(gpu)); }

  // And tell the algorithm how to set the QR parameters (BAFunctor.cpp:64-78)
  void initQRSolver(SchurlikeQRSolver& qr) {
    qr.h = gpu;
#if defined(QRKIT) || defined(QRCHOL)
    qr.setSparseBlockParams(numMeasurements * 2 + numPointParams, numPointParams);
#elif defined(MOREQR)
    qr.setSparseBlockParams(numMeasurements * 2, numPointParams);
#endif
  }
  void initQRSolverInner(SchurlikeQRSolver& qr) {
    qr.h = gpu;
#ifdef MOREQR
    qr.setSparseBlockParams(numPointParams * 2, numPointParams);
#endif
  }
};
#endif

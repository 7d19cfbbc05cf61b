// BAL problem-file CLI kept from the reference (src/bundle_adjustment_large.cpp:40-176): same usage
// string, return codes, loader conventions (negative focal, Rodrigues at load, k1*f^2, k2*f^4),
// before/after statistics (src/Utils.h:15-68) and per-variant LM dispatch by preprocessor symbol;
// the optimisation itself runs on the GPU behind include/ba_gpu.h.
//   Bundle_Adjustment_QRChol data/problem-21-11315-pre.txt
// Extra, optional: env BA_MAX_ITERS=<n> caps outer iterations, BA_LOG_CSV=<file> dumps the trial log,
// BA_DEVICE=<ordinal>.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "BATypeUtils.h"
#include "BAFunctor.h"
#include "BacktrackLevMarqGPU.h"

enum ReturnCodes { Success = 0, WrongInputParams = 1, WrongInputFile = 2 };

typedef BAFunctor OptimizationFunctor;

const Scalar AVG_FOCAL_LENGTH = 1.0;
const Scalar INLIER_THRESHOLD = 0.5;

namespace Math {
// src/MathUtils.h:66-82 (hard |omega| > 1e-6 cut-off)
inline void createRotationMatrixRodrigues(const double om[3], double R[9]) {
  const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (std::abs(theta) > 1e-6) {
    const double J[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double J2[9];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) J2[3 * a + b] = J[3 * a] * J[b] + J[3 * a + 1] * J[3 + b] + J[3 * a + 2] * J[6 + b];
    const double c1 = std::sin(theta) / theta, c2 = (1.0 - std::cos(theta)) / (theta * theta);
    for (int i = 0; i < 9; ++i) R[i] += c1 * J[i] + c2 * J2[i];
  }
}
}  // namespace Math

namespace Utils {
// src/Utils.h:15-68 as ONE reduction on the GPU at the functor's device-resident state (ba_error_statistics): the two calls
// print exactly the reference's three lines
inline double showErrorStatisticsAndObjective(OptimizationFunctor& functor, double avg_f, double inlierThreshold) {
  double s[4];
  ba_check(ba_error_statistics(functor.gpu, avg_f, inlierThreshold, s));
  const long K = functor.numMeasurements;
  const long nInliers = (long)s[2];
  std::cout << "Mean reprojection error: " << s[0] / K << std::endl;
  std::cout << "Inlier mean reprojection error: " << s[1] / nInliers << " (" << nInliers << " / " << K << " inliers)" << std::endl;
  std::cout << "True objective: " << s[3] << std::endl;
  return double(nInliers) / K;
}
}  // namespace Utils

// fast whitespace-separated number reader (the reference's ifstream>> parse takes seconds at 5M observations)
struct Tokens {
  std::vector<char> buf; const char* p; const char* end;
  explicit Tokens(const char* path) {
    std::ifstream ifs(path, std::ios::binary | std::ios::ate);
    if (!ifs) { p = end = nullptr; return; }
    const std::streamsize n = ifs.tellg(); ifs.seekg(0);
    buf.resize((size_t)n + 1); ifs.read(buf.data(), n); buf[(size_t)n] = 0;
    p = buf.data(); end = p + n;
  }
  bool ok() const { return p != nullptr; }
  double next() { char* q; const double v = std::strtod(p, &q); if (q == p) throw std::runtime_error("unexpected end of BAL file"); p = q; return v; }
};

int main(int argc, char* argv[]) {
  if (argc != 2) {
    std::cerr << "Usage: " << argv[0] << " <sparse reconstruction file>" << std::endl;
    return ReturnCodes::WrongInputParams;
  }
  Tokens tk(argv[1]);
  if (!tk.ok()) {
    std::cerr << "Cannot open " << argv[1] << std::endl;
    return ReturnCodes::WrongInputFile;
  }
  const double avg_focal_length = AVG_FOCAL_LENGTH;
  const int N = (int)tk.next(), M = (int)tk.next(), K = (int)tk.next();
  std::cout << "N(cameras) = " << N << ", M(points) = " << M << ", K(measurements) = " << K << std::endl;

  std::cout << "Reading image measurements..." << std::endl;
  std::vector<double> measurements(2 * (size_t)K);
  std::vector<int> correspondingView(K, -1), correspondingPoint(K, -1);
  for (int k = 0; k < K; ++k) {
    correspondingView[k] = (int)tk.next();
    correspondingPoint[k] = (int)tk.next();
    measurements[2 * (size_t)k] = tk.next() / avg_focal_length;
    measurements[2 * (size_t)k + 1] = tk.next() / avg_focal_length;
  }
  std::cout << "Done." << std::endl;

  std::cout << "Reading cameras params..." << std::endl;
  OptimizationFunctor::InputType params;
  params.R.resize(9 * (size_t)N); params.T.resize(3 * (size_t)N); params.f.resize(N); params.k1.resize(N); params.k2.resize(N);
  for (int i = 0; i < N; ++i) {
    double om[3], f, k1, k2;
    for (int r = 0; r < 3; ++r) om[r] = tk.next();
    for (int r = 0; r < 3; ++r) params.T[3 * (size_t)i + r] = tk.next();
    f = tk.next(); k1 = tk.next(); k2 = tk.next();
    params.f[i] = -f / avg_focal_length;                    // K(0,0) = K(1,1) = -f (:88-90)
    Math::createRotationMatrixRodrigues(om, &params.R[9 * (size_t)i]);
    const double f2 = f * f;
    params.k1[i] = k1 * f2; params.k2[i] = k2 * f2 * f2;    // :97-98
  }
  std::cout << "Done." << std::endl;

  std::cout << "Reading 3D points..." << std::endl;
  params.X.resize(3 * (size_t)M);
  for (size_t j = 0; j < 3 * (size_t)M; ++j) params.X[j] = tk.next();
  std::cout << "Done." << std::endl;

  // the kernels need observations grouped by point (the reference's row permutation relies on it): stable sort
  {
    std::vector<int> order(K);
    std::iota(order.begin(), order.end(), 0);
    bool sorted = true;
    for (int k = 1; k < K && sorted; ++k)
      sorted = correspondingPoint[k] > correspondingPoint[k - 1] || (correspondingPoint[k] == correspondingPoint[k - 1] && correspondingView[k] > correspondingView[k - 1]);
    if (!sorted) {
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return correspondingPoint[a] != correspondingPoint[b] ? correspondingPoint[a] < correspondingPoint[b] : correspondingView[a] < correspondingView[b]; });
      std::vector<int> v(K), p(K); std::vector<double> m(2 * (size_t)K);
      for (int k = 0; k < K; ++k) { v[k] = correspondingView[order[k]]; p[k] = correspondingPoint[order[k]]; m[2 * (size_t)k] = measurements[2 * (size_t)order[k]]; m[2 * (size_t)k + 1] = measurements[2 * (size_t)order[k] + 1]; }
      correspondingView.swap(v); correspondingPoint.swap(p); measurements.swap(m);
    }
  }

  const int device = std::getenv("BA_DEVICE") ? std::atoi(std::getenv("BA_DEVICE")) : 0;
  try {
    OptimizationFunctor functor(M, N, measurements, correspondingView, correspondingPoint, INLIER_THRESHOLD, device);
    functor.upload(params);
    Utils::showErrorStatisticsAndObjective(functor, avg_focal_length, INLIER_THRESHOLD);   // :130-131
    BacktrackLevMarqGPU<OptimizationFunctor, true> lm(functor);
    if (std::getenv("BA_MAX_ITERS")) lm.lmParams().maxIter = std::atoi(std::getenv("BA_MAX_ITERS"));
    const auto begin = std::chrono::steady_clock::now();
    BacktrackLevMarqInfo::Status info = lm.minimize(params);
    std::cout << "lm.minimize(params) ... " << std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count() << "s" << std::endl;
    std::cout << "LM finished with status: " << BacktrackLevMarqInfo::statusToString(info) << std::endl;
    if (const char* csv = std::getenv("BA_LOG_CSV")) {
      std::ofstream o(csv);
      o.precision(17);
      o << "iter,accepted,energy,energy_test,rho,lambda_used,lambda_next,dx_norm,elapsed_s\n";
      for (const auto& t : lm.log()) o << t.iter << ',' << t.accepted << ',' << t.energy << ',' << t.energyTest << ',' << t.rho << ',' << t.lambdaUsed << ',' << t.lambdaNext << ',' << t.dxNorm << ',' << t.elapsed << '\n';
    }
    Utils::showErrorStatisticsAndObjective(functor, avg_focal_length, INLIER_THRESHOLD);   // :170-171, at the final state on the device
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 3;
  }

  return ReturnCodes::Success;
}

// LM control flow kept on the host, one template for the four variant symbols. Mirrors
// src/Eigen_ext/BacktrackLevMarqQRChol.h:204-436 (QRCHOL; QRKIT's own loop is not in the tree and is
// assumed identical), BacktrackLevMarqMore.h:204-425 (MOREQR: lambda_0 = 1e-6*max||J(:,c)||) and
// BacktrackLevMarqCholesky.h:190-361 (CHOLESKY). Same constants, same iteration table.
#ifndef BACKTRACK_LEVMARQ_GPU_H
#define BACKTRACK_LEVMARQ_GPU_H
#include <algorithm>
#include <chrono>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "BAFunctor.h"

namespace BacktrackLevMarqInfo {
enum Status { NotStarted = -2, Running = -1, Success = 0, ExceededLambdaMax = 1, TooManyFunctionEvaluation = 2, MaxItersReached = 3 };

inline std::string statusToString(const Status& status) {
  switch (status) {
    case NotStarted: return "Not Started";
    case Running: return "Running";
    case Success: return "Success (Energy Flatlined)";
    case ExceededLambdaMax: return "Success (Exceeded Maximum Lambda)";
    case TooManyFunctionEvaluation: return "Too Many Function Evaluations";
    case MaxItersReached: return "Maximum Iterations Reached";
  }
  return "";
}
inline void outputHeader() {
  std::cout << "############################## Backtrack LevMarq ###############################" << std::endl;
  std::cout << "--------------------------------------------------------------------------------" << std::endl;
}
inline void outputFooter() { std::cout << "--------------------------------------------------------------------------------" << std::endl; }
inline void outputIterHeader() {
  std::cout << " Iter" << std::setw(15) << "Status" << std::setw(15) << "f" << std::setw(15) << "rho" << std::setw(15) << "lambda"
            << std::setw(15) << "Elapsed" << std::endl;
  std::cout << "--------------------------------------------------------------------------------" << std::endl;
}
template <typename Index, typename S>
void outputIter(const Index iterIdx, const std::string& status, const S& fVal, const S& rho, const S& lambda, const double& elapsed) {
  std::cout << std::setw(5) << iterIdx << std::setw(15) << status << std::setw(15) << fVal << std::setw(15) << rho << std::setw(15)
            << lambda << std::setw(14) << elapsed << "s" << std::endl;
}
}  // namespace BacktrackLevMarqInfo

struct TrialRecord { int iter; bool accepted; double energy, energyTest, rho, lambdaUsed, lambdaNext, dxNorm, elapsed; };

template <typename _FunctorType, bool Verbose = false>
class BacktrackLevMarqGPU {
 public:
  typedef _FunctorType FunctorType;
  typedef typename FunctorType::QRSolver QRSolver;
  typedef typename FunctorType::InputType InputType;

  struct Lambda { Scalar minVal, maxVal, decrease, increaseBase, init; Lambda() : minVal(1e-10), maxVal(1e10), decrease(10), increaseBase(2), init(1e-3) {} };
  struct LMParams { Lambda lambda; Scalar tolFun; int maxIter; int maxFunEv; LMParams() : lambda(Lambda()), tolFun(1e-8), maxIter(1e6), maxFunEv(1e6) {} };
  struct OptimParams {
    Scalar lambda, lambdaIncrease; int funEvals, iter;
    void initialize(LMParams& p) { lambda = p.lambda.init; lambdaIncrease = p.lambda.increaseBase; funEvals = 0; iter = 0; }
  };
  const int EnergyHistorySize = 2;

  explicit BacktrackLevMarqGPU(FunctorType& functor) : m_functor(functor), m_status(BacktrackLevMarqInfo::NotStarted) {}
  LMParams& lmParams() { return m_lmParams; }
  const std::vector<TrialRecord>& log() const { return m_log; }

  BacktrackLevMarqInfo::Status minimize(InputType& x) {
    using namespace BacktrackLevMarqInfo;
    typedef std::chrono::steady_clock Clock;
    if (Verbose) outputHeader();
    m_optParams.initialize(m_lmParams);
    m_functor.initQRSolver(m_solver);
    m_functor.initQRSolverInner(m_solver);
    m_functor.upload(x);
    m_energyHistory.assign(EnergyHistorySize, Scalar(0));
    m_status = Running;
    bool stopNow = false;
    if (Verbose) outputIterHeader();
    Clock::time_point iterStart;
    while (true) {
      iterStart = Clock::now();
      m_optParams.iter++;
      if (m_optParams.iter > m_lmParams.maxIter) { m_status = MaxItersReached; break; }
      if (m_optParams.funEvals > m_lmParams.maxFunEv) { m_status = TooManyFunctionEvaluation; break; }
      // r = f(x); energy; J = df(x); JtRes; column norms   (QRChol.h:257-280)
      Scalar maxCn2 = 0, maxCn = 0;
      const bool first = m_optParams.iter == 1;
      m_functor.df(m_energy, first ? &maxCn2 : nullptr, first ? &maxCn : nullptr);
      m_optParams.funEvals++;
      if (first) {
#if defined(MOREQR)
        m_optParams.lambda = 1e-6 * maxCn;    // More.h:284
#else
        m_optParams.lambda = 1e-12 * maxCn2;  // QRChol.h:279, Cholesky.h:264
#endif
      }
      while (true) {
#if !defined(MOREQR)
        iterStart = Clock::now();  // QRChol.h:284 / Cholesky.h:269 restart the timer per trial
#endif
        const Scalar lambdaUsed = m_optParams.lambda;
        m_solver.compute(m_optParams.lambda);
        Scalar dxNorm, rhoScale, energyTest;
        m_solver.solve(dxNorm, rhoScale, energyTest);
        m_optParams.funEvals++;
        const double elapsed = std::chrono::duration<double>(Clock::now() - iterStart).count();
        if (energyTest < m_energy) {
          Scalar rho = (m_energy - energyTest) / rhoScale;
          Scalar lambdaMul = Scalar(1.0) - std::pow(Scalar(2.0) * rho - Scalar(1.0), Scalar(3.0));
          m_optParams.lambda *= std::max<Scalar>(Scalar(1.0) / Scalar(3.0), lambdaMul);
          m_optParams.lambda = std::max<Scalar>(m_optParams.lambda, m_lmParams.lambda.minVal);
          if (Verbose) outputIter<int, Scalar>(m_optParams.iter, "Accepted", m_energy, rho, m_optParams.lambda, elapsed);
          m_log.push_back({m_optParams.iter, true, (double)m_energy, (double)energyTest, (double)rho, (double)lambdaUsed, (double)m_optParams.lambda, (double)dxNorm, elapsed});
          m_optParams.lambdaIncrease = m_lmParams.lambda.increaseBase;
          m_energy = energyTest;
          m_energyHistory[m_optParams.iter % EnergyHistorySize] = m_energy;
          break;
        } else {
          if (Verbose) outputIter<int, Scalar>(m_optParams.iter, "Rejected", m_energy, Scalar(0), m_optParams.lambda, elapsed);
          m_log.push_back({m_optParams.iter, false, (double)m_energy, (double)energyTest, 0.0, (double)lambdaUsed, (double)m_optParams.lambda, (double)dxNorm, elapsed});
          m_functor.reject();
#if defined(MOREQR)
          iterStart = Clock::now();  // More.h:386
#endif
          if (m_optParams.lambda > m_lmParams.lambda.maxVal) { m_status = ExceededLambdaMax; stopNow = true; break; }
          m_optParams.lambda *= m_optParams.lambdaIncrease;
#if defined(QRCHOL) || defined(QRKIT)
          m_optParams.lambdaIncrease = std::pow(m_optParams.lambdaIncrease, 1.5f);  // QRChol.h:409 (float literal, quirk Q5)
#else
          m_optParams.lambdaIncrease = std::pow(m_optParams.lambdaIncrease, 1.5);
#endif
        }
      }
      if (stopNow) break;
      if (m_optParams.iter > EnergyHistorySize) {
        Scalar maxf = *(std::max_element(m_energyHistory.begin(), m_energyHistory.end()));
        if (std::abs(m_energy - maxf) < m_lmParams.tolFun * m_energy) { m_status = Success; break; }  // x not committed (quirk Q8)
      }
      m_functor.accept();  // x = xTest (QRChol.h:428)
    }
    m_functor.download(x);
    if (Verbose) outputFooter();
    return m_status;
  }

 private:
  FunctorType& m_functor;
  QRSolver m_solver;
  Scalar m_energy = 0;
  std::vector<Scalar> m_energyHistory;
  LMParams m_lmParams;
  OptimParams m_optParams;
  BacktrackLevMarqInfo::Status m_status;
  std::vector<TrialRecord> m_log;
};
#endif

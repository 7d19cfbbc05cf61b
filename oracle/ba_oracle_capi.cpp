// TEST INFRASTRUCTURE — NOT PRODUCT CODE. extern "C" surface of the CPU oracle (ba_oracle.hpp) for
// ctypes. All scalars cross this boundary as double; `precision` (0 = float, 1 = double) selects the
// Scalar the arithmetic runs in (reference: src/BATypeUtils.h:6-7).
#include "ba_oracle.hpp"

namespace {
struct Base {
  virtual ~Base() {}
  virtual void set_state(const double*, const double*, const double*, const double*, const double*, const double*) = 0;
  virtual void get_state(double*, double*, double*, double*, double*, double*) = 0;
  virtual void linearize(double*, double*, double*) = 0;
  virtual void residuals(double*) = 0;
  virtual void jacobian(double*, double*) = 0;
  virtual void jtres(double*) = 0;
  virtual void moreqr_outer() = 0;
  virtual int step(int, double, double*) = 0;
  virtual double energy_at(const double*) = 0;
  virtual void apply(const double*) = 0;
  virtual void reduced(double*, double*) = 0;
  virtual void set_tall(int) = 0;
  virtual void set_threads(int) = 0;
  virtual int kd() = 0;
  virtual int minimize(int, int, bao::TrialRecord*, int, int*) = 0;
};
template <class S>
struct Impl : Base {
  bao::Oracle<S> o;
  template <class A, class B> static void cp(const A* a, std::vector<B>& b) { for (size_t i = 0; i < b.size(); ++i) b[i] = B(a[i]); }
  template <class A, class B> static void cpo(const std::vector<A>& a, B* b) { for (size_t i = 0; i < a.size(); ++i) b[i] = B(a[i]); }
  void set_state(const double* R, const double* T, const double* f, const double* k1, const double* k2, const double* X) override {
    cp(R, o.R); cp(T, o.T); cp(f, o.f); cp(k1, o.k1); cp(k2, o.k2); cp(X, o.X);
  }
  void get_state(double* R, double* T, double* f, double* k1, double* k2, double* X) override {
    cpo(o.R, R); cpo(o.T, T); cpo(o.f, f); cpo(o.k1, k1); cpo(o.k2, k2); cpo(o.X, X);
  }
  void linearize(double* e, double* a, double* b) override { o.linearize(e, a, b); }
  void residuals(double* r) override { std::vector<S> fv; o.residuals(o.R, o.T, o.f, o.k1, o.k2, o.X, fv); cpo(fv, r); }
  void jacobian(double* Jc, double* Jp) override { cpo(o.Jc, Jc); cpo(o.Jp, Jp); }
  void jtres(double* v) override { cpo(o.JtRes, v); }
  void moreqr_outer() override { o.moreqr_outer(); }
  int step(int v, double lam, double* dx) override {
    std::vector<S> d; bool ok = o.step((bao::Variant)v, lam, d); cpo(d, dx); return ok ? 1 : 0;
  }
  double energy_at(const double* dx) override { std::vector<S> d(o.nparams()); cp(dx, d); return o.energy_at(d); }
  void apply(const double* dx) override { std::vector<S> d(o.nparams()); cp(dx, d); o.apply(d); }
  void reduced(double* Sd, double* g) override {  // dense symmetric 9N x 9N row-major, and g
    const int n = 9 * o.N, kd = o.kd;
    for (size_t i = 0; i < (size_t)n * n; ++i) Sd[i] = 0;
    for (int i = 0; i < n; ++i) for (int j = std::max(0, i - kd); j <= i; ++j) {
      const double v = (double)o.S_last[(size_t)i * (kd + 1) + (j - i + kd)];
      Sd[(size_t)i * n + j] = v; Sd[(size_t)j * n + i] = v;
    }
    cpo(o.g_last, g);
  }
  void set_tall(int t) override { o.tall_qr = t != 0; }
  void set_threads(int t) override { o.nthreads = t < 1 ? 1 : t; }
  int kd() override { return o.kd; }
  int minimize(int v, int max_outer, bao::TrialRecord* log, int cap, int* nlog) override {
    std::vector<bao::TrialRecord> l;
    int st = (int)o.minimize((bao::Variant)v, max_outer, l);
    int n = (int)std::min<size_t>(l.size(), (size_t)cap);
    for (int i = 0; i < n; ++i) log[i] = l[i];
    *nlog = (int)l.size();
    return st;
  }
};
}  // namespace

extern "C" {
void* bao_create(int precision, int N, int M, int K, const int* view, const int* point, const double* meas, double tau) {
  if (precision == 0) { auto* p = new Impl<float>(); p->o.init(N, M, K, view, point, meas, tau); return p; }
  auto* p = new Impl<double>(); p->o.init(N, M, K, view, point, meas, tau); return p;
}
void bao_destroy(void* h) { delete (Base*)h; }
void bao_set_state(void* h, const double* R, const double* T, const double* f, const double* k1, const double* k2, const double* X) { ((Base*)h)->set_state(R, T, f, k1, k2, X); }
void bao_get_state(void* h, double* R, double* T, double* f, double* k1, double* k2, double* X) { ((Base*)h)->get_state(R, T, f, k1, k2, X); }
void bao_linearize(void* h, double* energy, double* max_colnorm2, double* max_colnorm) { ((Base*)h)->linearize(energy, max_colnorm2, max_colnorm); }
void bao_residuals(void* h, double* r) { ((Base*)h)->residuals(r); }
void bao_jacobian(void* h, double* Jc, double* Jp) { ((Base*)h)->jacobian(Jc, Jp); }
void bao_jtres(void* h, double* v) { ((Base*)h)->jtres(v); }
void bao_moreqr_outer(void* h) { ((Base*)h)->moreqr_outer(); }
int bao_step(void* h, int variant, double lambda, double* dx) { return ((Base*)h)->step(variant, lambda, dx); }
double bao_energy_at(void* h, const double* dx) { return ((Base*)h)->energy_at(dx); }
void bao_apply(void* h, const double* dx) { ((Base*)h)->apply(dx); }
void bao_reduced(void* h, double* S, double* g) { ((Base*)h)->reduced(S, g); }
void bao_set_tall(void* h, int t) { ((Base*)h)->set_tall(t); }
int bao_has_openmp(void) {
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}
void bao_set_threads(void* h, int t) { ((Base*)h)->set_threads(t); }
int bao_kd(void* h) { return ((Base*)h)->kd(); }
int bao_minimize(void* h, int variant, int max_outer, bao::TrialRecord* log, int cap, int* nlog) { return ((Base*)h)->minimize(variant, max_outer, log, cap, nlog); }
}

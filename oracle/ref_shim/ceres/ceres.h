// Minimal stand-in for <ceres/ceres.h>, just enough to compile and EVALUATE the
// REFERENCE's own functors in DCS-ceres/src/ceres_error.cpp where they lie
// (oracle/Makefile target `ref`).  TEST INFRASTRUCTURE ONLY.
// Ceres Solver is an un-vendored, un-versioned dependency of the reference
// (DCS-ceres/CMakeLists.txt:9), absent offline.  Restated here from its published
// semantics: Jet<T,N> arithmetic (jet.h rules) and AutoDiffCostFunction::Evaluate
// (seed parameter i of block b with e_{offset_b+i}; jacobians[b] row-major
// kNumResiduals x N_b).  The minimiser is NOT restated here (see oracle/dcs_oracle.cpp).
#ifndef DCS_REF_SHIM_CERES_H
#define DCS_REF_SHIM_CERES_H
#include <cmath>

namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& s) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); }  // NOLINT (implicit, like Ceres)
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1); }
};
#define DCS_JET template <typename T, int N> inline
DCS_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
DCS_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
DCS_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
DCS_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
DCS_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  const T g_a_inverse = T(1.0) / g.a;
  const T f_a_by_g_a = f.a * g_a_inverse;
  Jet<T, N> h; h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
DCS_JET bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
DCS_JET Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
DCS_JET Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T s = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
DCS_JET Jet<T, N> asin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::asin(f.a); const T t = T(1.0) / std::sqrt(T(1.0) - f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
DCS_JET Jet<T, N> sqrt(const Jet<T, N>& f) { Jet<T, N> h; const T t = std::sqrt(f.a); h.a = t; const T two_a_inverse = T(1.0) / (T(2.0) * t); for (int i = 0; i < N; ++i) h.v[i] = two_a_inverse * f.v[i]; return h; }
#undef DCS_JET

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};

class LossFunction;

template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
  static constexpr int kBlocks = sizeof...(Ns);
  static constexpr int kTotal = (Ns + ...);

 public:
  explicit AutoDiffCostFunction(Functor* f) : f_(f) {}
  ~AutoDiffCostFunction() override { delete f_; }

  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    const int sizes[kBlocks] = {Ns...};
    if (!jacobians) return call<double>(parameters, residuals, sizes);
    typedef Jet<double, kTotal> J;
    J x[kTotal];
    const J* ptr[kBlocks];
    int off = 0;
    for (int b = 0; b < kBlocks; ++b) {
      ptr[b] = x + off;
      for (int i = 0; i < sizes[b]; ++i) x[off + i] = J(parameters[b][i], off + i);
      off += sizes[b];
    }
    J e[kNumResiduals];
    if (!invoke(ptr, e, std::make_integer_sequence<int, kBlocks>())) return false;
    off = 0;
    for (int b = 0; b < kBlocks; ++b) {
      if (jacobians[b])
        for (int r = 0; r < kNumResiduals; ++r)
          for (int i = 0; i < sizes[b]; ++i) jacobians[b][r * sizes[b] + i] = e[r].v[off + i];
      off += sizes[b];
    }
    for (int r = 0; r < kNumResiduals; ++r) residuals[r] = e[r].a;
    return true;
  }

 private:
  template <typename T>
  bool call(double const* const* parameters, double* residuals, const int*) const {
    return invoke(parameters, residuals, std::make_integer_sequence<int, kBlocks>());
  }
  template <typename T, int... I>
  bool invoke(T const* const* p, T* e, std::integer_sequence<int, I...>) const {
    return (*f_)(p[I]..., e);
  }
  Functor* f_;
};

}  // namespace ceres
#endif

// dcs_oracle.cpp — CPU restatement of the reference's DCS-ceres solve path.
//
// TEST INFRASTRUCTURE ONLY (see dcs_oracle.h).  Not linked into, loaded by or called from
// the product library; it is the checker in tests/ and the timed CPU baseline in bench.py.
//
// What is restated, and from where (paths relative to the reference checkout):
//   * the two residual functors, literally, evaluated on a hand-written Jet<6> exactly as
//     ceres::AutoDiffCostFunction<F,3,3,3> would:      DCS-ceres/src/ceres_error.cpp:42-94
//     (OdometryResidue::operator()) and :135-196 (DCSClosureResidue::operator());
//     measurement matrix built as in the constructors   :4-25 / :97-118
//   * Eigen's fixed-size 3x3 inverse (cofactors, det along column 0, multiply by 1/det)
//     and coefficient-wise 3x3 product — third-party (Eigen3, version unpinned by
//     DCS-ceres/CMakeLists.txt:8), restated from its published algorithm
//   * ceres::HuberLoss(0.01) + Corrector on every block: DCS-ceres/main.cpp:67-68 and the
//     AddResidualBlock call sites :99,:114,:128,:137,:148
//   * residual-block order / constant pose 0:            DCS-ceres/main.cpp:95-150,:153
//   * Ceres-default trust-region Levenberg–Marquardt with SPARSE_NORMAL_CHOLESKY
//     (DCS-ceres/main.cpp:154-163) — third-party (Ceres Solver, version unpinned by
//     DCS-ceres/CMakeLists.txt:9, absent offline): restated from Ceres 2.x
//     trust_region_minimizer.cc / levenberg_marquardt_strategy.cc / loss_function.cc /
//     corrector.cc semantics.  PARITY UNPINNED for this part.
//
// Build: see oracle/Makefile (g++ -O2 -fopenmp, no dependencies).

#include "dcs_oracle.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------
// Jet<double,6>: value + 6 partials, rules as in Ceres' jet.h.
// ------------------------------------------------------------------------------------
struct Jet {
  double a;
  double v[6];
  Jet() : a(0.0) { for (double& x : v) x = 0.0; }
  explicit Jet(double s) : a(s) { for (double& x : v) x = 0.0; }
  Jet(double s, int k) : a(s) { for (double& x : v) x = 0.0; v[k] = 1.0; }
};
inline Jet operator+(const Jet& f, const Jet& g) { Jet h; h.a = f.a + g.a; for (int i = 0; i < 6; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
inline Jet operator-(const Jet& f, const Jet& g) { Jet h; h.a = f.a - g.a; for (int i = 0; i < 6; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
inline Jet operator-(const Jet& f) { Jet h; h.a = -f.a; for (int i = 0; i < 6; ++i) h.v[i] = -f.v[i]; return h; }
inline Jet operator*(const Jet& f, const Jet& g) { Jet h; h.a = f.a * g.a; for (int i = 0; i < 6; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
inline Jet operator/(const Jet& f, const Jet& g) {
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  Jet h; h.a = f_a_by_g_a;
  for (int i = 0; i < 6; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
inline bool operator<(const Jet& f, const Jet& g) { return f.a < g.a; }
inline Jet sin(const Jet& f) { Jet h; h.a = std::sin(f.a); const double c = std::cos(f.a); for (int i = 0; i < 6; ++i) h.v[i] = c * f.v[i]; return h; }
inline Jet cos(const Jet& f) { Jet h; h.a = std::cos(f.a); const double s = -std::sin(f.a); for (int i = 0; i < 6; ++i) h.v[i] = s * f.v[i]; return h; }
inline Jet asin(const Jet& f) { Jet h; h.a = std::asin(f.a); const double t = 1.0 / std::sqrt(1.0 - f.a * f.a); for (int i = 0; i < 6; ++i) h.v[i] = t * f.v[i]; return h; }
inline Jet sqrt(const Jet& f) { Jet h; const double t = std::sqrt(f.a); h.a = t; const double two_a_inverse = 1.0 / (2.0 * t); for (int i = 0; i < 6; ++i) h.v[i] = two_a_inverse * f.v[i]; return h; }

template <typename T> inline T make(double s);
template <> inline double make<double>(double s) { return s; }
template <> inline Jet make<Jet>(double s) { return Jet(s); }
template <typename T> inline T tmin(const T& a, const T& b) { return (b < a) ? b : a; }  // std::min

// ------------------------------------------------------------------------------------
// 3x3 matrix with Eigen's fixed-size inverse and coefficient-based product.
// ------------------------------------------------------------------------------------
template <typename T> struct M3 {
  T m[3][3];
  T& operator()(int i, int j) { return m[i][j]; }
  const T& operator()(int i, int j) const { return m[i][j]; }
};
template <typename T> inline T cof(const M3<T>& m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
}
template <typename T> inline M3<T> inverse(const M3<T>& m) {
  M3<T> r;
  const T c0 = cof(m, 0, 0), c1 = cof(m, 1, 0), c2 = cof(m, 2, 0);
  const T det = (c0 * m(0, 0) + c1 * m(1, 0)) + c2 * m(2, 0);
  const T invdet = make<T>(1.0) / det;
  const T c01 = cof(m, 0, 1) * invdet;
  const T c11 = cof(m, 1, 1) * invdet;
  const T c02 = cof(m, 0, 2) * invdet;
  r(1, 2) = cof(m, 2, 1) * invdet;
  r(2, 1) = cof(m, 1, 2) * invdet;
  r(2, 2) = cof(m, 2, 2) * invdet;
  r(1, 0) = c01;
  r(1, 1) = c11;
  r(2, 0) = c02;
  r(0, 0) = c0 * invdet;
  r(0, 1) = c1 * invdet;
  r(0, 2) = c2 * invdet;
  return r;
}
template <typename T> inline M3<T> mul(const M3<T>& a, const M3<T>& b) {
  M3<T> r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
  return r;
}

// ------------------------------------------------------------------------------------
// The functors (ceres_error.cpp:42-94 plain, :135-196 DCS).
// ------------------------------------------------------------------------------------
struct Meas { double M[3][3]; };
inline Meas make_meas(double dx, double dy, double dth) {  // ctor :4-25 / :97-118
  Meas q;
  const double c = std::cos(dth), s = std::sin(dth);
  q.M[0][0] = c; q.M[0][1] = -s; q.M[1][0] = s; q.M[1][1] = c;
  q.M[0][2] = dx; q.M[1][2] = dy; q.M[2][0] = 0.0; q.M[2][1] = 0.0; q.M[2][2] = 1.0;
  return q;
}
using std::sin; using std::cos; using std::asin; using std::sqrt;
template <typename T> inline M3<T> pose_matrix(const T* P) {
  M3<T> w;
  const T c = cos(P[2]);
  const T s = sin(P[2]);
  w(0, 0) = c; w(0, 1) = -s; w(1, 0) = s; w(1, 1) = c;
  w(0, 2) = P[0]; w(1, 2) = P[1];
  w(2, 0) = make<T>(0.0); w(2, 1) = make<T>(0.0); w(2, 2) = make<T>(1.0);
  return w;
}
template <typename T>
inline void functor(const Meas& q, bool dcs, double phi, const T* P1, const T* P2, T* e, T* psi_out) {
  const M3<T> w_T_a = pose_matrix(P1);
  const M3<T> w_T_b = pose_matrix(P2);
  M3<T> Tm;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Tm(i, j) = make<T>(q.M[i][j]);
  const M3<T> diff = mul(inverse(Tm), mul(inverse(w_T_a), w_T_b));   // :87 / :180
  if (!dcs) {
    e[0] = diff(0, 2);
    e[1] = diff(1, 2);
    e[2] = asin(diff(1, 0));
    *psi_out = make<T>(1.0);
    return;
  }
  const T res = diff(0, 2) * diff(0, 2) + diff(1, 2) * diff(1, 2);                 // :186
  const T psi_org = sqrt(make<T>(2.0) * make<T>(phi) / (make<T>(phi) + res));      // :187
  const T psi = tmin(make<T>(1.0), psi_org);                                       // :188
  e[0] = psi * diff(0, 2);
  e[1] = psi * diff(1, 2);
  e[2] = psi * asin(diff(1, 0));
  *psi_out = psi;
}

struct Huber {  // ceres::HuberLoss(a)
  double a, b;
  explicit Huber(double a_) : a(a_), b(a_ * a_) {}
  void eval(double s, double rho[3]) const {
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
};

inline bool edge_uses_dcs(const oracle_problem* p, int k) {  // main.cpp:55,112,135
  return p->dcs_on && p->kind[k] != 0;
}

// Per-edge: AutoDiff evaluation (+ optional corrector). r[3], J[18] row-major 3x6.
inline void eval_edge_jet(const oracle_problem* p, const double* x, int k, bool raw,
                          double* r, double* J, double* psi, double* rho1, double* cost) {
  const int ia = p->edge_a[k], ib = p->edge_b[k];
  const Meas q = make_meas(p->meas_xyt[3 * k], p->meas_xyt[3 * k + 1], p->meas_xyt[3 * k + 2]);
  Jet P1[3], P2[3], e[3], ps;
  for (int i = 0; i < 3; ++i) { P1[i] = Jet(x[3 * ia + i], i); P2[i] = Jet(x[3 * ib + i], 3 + i); }
  functor<Jet>(q, edge_uses_dcs(p, k), p->phi, P1, P2, e, &ps);
  for (int i = 0; i < 3; ++i) { r[i] = e[i].a; for (int j = 0; j < 6; ++j) J[6 * i + j] = e[i].v[j]; }
  const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  double rho[3];
  Huber(p->huber_delta).eval(s, rho);
  *cost = 0.5 * rho[0];
  *psi = ps.a;
  *rho1 = rho[1];
  if (!raw) {
    // Corrector: s == 0 or rho[2] <= 0 always holds for Huber -> scale by sqrt(rho').
    const double sc = std::sqrt(rho[1]);
    for (int i = 0; i < 18; ++i) J[i] *= sc;
    for (int i = 0; i < 3; ++i) r[i] *= sc;
  }
}

inline double cost_edge_double(const oracle_problem* p, const double* x, int k) {
  const int ia = p->edge_a[k], ib = p->edge_b[k];
  const Meas q = make_meas(p->meas_xyt[3 * k], p->meas_xyt[3 * k + 1], p->meas_xyt[3 * k + 2]);
  double e[3], ps;
  functor<double>(q, edge_uses_dcs(p, k), p->phi, x + 3 * ia, x + 3 * ib, e, &ps);
  const double s = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  double rho[3];
  Huber(p->huber_delta).eval(s, rho);
  return 0.5 * rho[0];
}

// METHOD 2 (switchable constraints): loop edges (kind >= 1) carry a scalar switch s, e = s * e_plain
// (ceres_error.cpp:238-297), Huber on that block; plus the prior row sqrt(lambda) (1 - s) without loss
// (:300-317, main.cpp:105-150).  Corrected residual r[3], pose Jacobian J[18] (3x6), switch Jacobian Js[3].
inline void eval_edge_jet_sc(const oracle_problem* p, const double* x, double s, int k,
                             double* r, double* J, double* Js, double* cost) {
  const int ia = p->edge_a[k], ib = p->edge_b[k];
  const Meas q = make_meas(p->meas_xyt[3 * k], p->meas_xyt[3 * k + 1], p->meas_xyt[3 * k + 2]);
  Jet P1[3], P2[3], d[3], ps;
  for (int i = 0; i < 3; ++i) { P1[i] = Jet(x[3 * ia + i], i); P2[i] = Jet(x[3 * ib + i], 3 + i); }
  functor<Jet>(q, false, p->phi, P1, P2, d, &ps);
  for (int i = 0; i < 3; ++i) {
    r[i] = s * d[i].a;
    for (int j = 0; j < 6; ++j) J[6 * i + j] = s * d[i].v[j] + 0.0 * d[i].a;
    Js[i] = s * 0.0 + 1.0 * d[i].a;
  }
  const double sq = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  double rho[3];
  Huber(p->huber_delta).eval(sq, rho);
  *cost = 0.5 * rho[0];
  const double sc = std::sqrt(rho[1]);
  for (int i = 0; i < 18; ++i) J[i] *= sc;
  for (int i = 0; i < 3; ++i) { r[i] *= sc; Js[i] *= sc; }
}
inline double cost_edge_double_sc(const oracle_problem* p, const double* x, double s, double lambda, int k) {
  const int ia = p->edge_a[k], ib = p->edge_b[k];
  const Meas q = make_meas(p->meas_xyt[3 * k], p->meas_xyt[3 * k + 1], p->meas_xyt[3 * k + 2]);
  double e[3], ps;
  functor<double>(q, false, p->phi, x + 3 * ia, x + 3 * ib, e, &ps);
  double rho[3];
  Huber(p->huber_delta).eval((s * e[0]) * (s * e[0]) + (s * e[1]) * (s * e[1]) + (s * e[2]) * (s * e[2]), rho);   // |s e|^2
  const double pr = std::sqrt(lambda) * (1.0 - s);
  return 0.5 * rho[0] + 0.5 * pr * pr;
}

// Per-loop-edge switch blocks of the normal equations (METHOD 2); empty for METHOD 0/1.
struct SwitchBlocks {
  bool on = false;
  double lambda = 1.0;
  std::vector<double> s;        // E switch values (entries of odometry edges unused)
  std::vector<double> hss, gs;  // J_s^T J_s + lambda,  J_s^T r - lambda (1 - s)
  std::vector<double> hps;      // E x 6: J_a^T J_s, J_b^T J_s
  bool has(const oracle_problem* p, int k) const { return on && p->kind[k] != 0; }
};

int set_threads(const oracle_problem* p) {
  int t = p->num_threads > 0 ? p->num_threads : 1;
#ifdef _OPENMP
  omp_set_num_threads(t);
#else
  t = 1;
#endif
  return t;
}

// ------------------------------------------------------------------------------------
// Block structure of J^T J.
// ------------------------------------------------------------------------------------
struct Structure {
  int N = 0, E = 0;
  std::vector<char> is_free;           // pose is a non-constant parameter touched by an edge
  std::vector<int> row_ptr, col_idx;   // upper block CSR over poses (diag first in each row)
  std::vector<int> slot_aa, slot_bb, slot_ab;  // per edge; -1 when an endpoint is not free
  std::vector<char> ab_transposed;     // 1 when a > b: block stored is (b,a) = B^T A
  // contributions per block, for the parallel deterministic accumulate
  std::vector<int> contrib_ptr, contrib;  // contrib = edge*4 + which (0 aa, 1 bb, 2 ab)
};

void build_structure(const oracle_problem* p, Structure& S) {
  const int N = p->n_poses, E = p->n_edges;
  S.N = N; S.E = E;
  S.is_free.assign(N, 0);
  for (int k = 0; k < E; ++k) { S.is_free[p->edge_a[k]] = 1; S.is_free[p->edge_b[k]] = 1; }
  if (p->fixed_pose >= 0 && p->fixed_pose < N) S.is_free[p->fixed_pose] = 0;
  std::vector<std::vector<int>> cols(N);
  for (int i = 0; i < N; ++i) if (S.is_free[i]) cols[i].push_back(i);
  for (int k = 0; k < E; ++k) {
    const int a = p->edge_a[k], b = p->edge_b[k];
    if (!S.is_free[a] || !S.is_free[b]) continue;
    cols[std::min(a, b)].push_back(std::max(a, b));
  }
  S.row_ptr.assign(N + 1, 0);
  S.col_idx.clear();
  for (int i = 0; i < N; ++i) {
    std::sort(cols[i].begin(), cols[i].end());
    cols[i].erase(std::unique(cols[i].begin(), cols[i].end()), cols[i].end());
    S.row_ptr[i] = (int)S.col_idx.size();
    S.col_idx.insert(S.col_idx.end(), cols[i].begin(), cols[i].end());
  }
  S.row_ptr[N] = (int)S.col_idx.size();
  auto find = [&](int r, int c) {
    const int* b = S.col_idx.data() + S.row_ptr[r];
    const int* e = S.col_idx.data() + S.row_ptr[r + 1];
    return (int)(std::lower_bound(b, e, c) - S.col_idx.data());
  };
  S.slot_aa.assign(E, -1); S.slot_bb.assign(E, -1); S.slot_ab.assign(E, -1); S.ab_transposed.assign(E, 0);
  const int nb = (int)S.col_idx.size();
  std::vector<int> cnt(nb + 1, 0);
  for (int k = 0; k < E; ++k) {
    const int a = p->edge_a[k], b = p->edge_b[k];
    if (S.is_free[a]) { S.slot_aa[k] = find(a, a); cnt[S.slot_aa[k] + 1]++; }
    if (S.is_free[b]) { S.slot_bb[k] = find(b, b); cnt[S.slot_bb[k] + 1]++; }
    if (S.is_free[a] && S.is_free[b]) {
      S.slot_ab[k] = find(std::min(a, b), std::max(a, b));
      S.ab_transposed[k] = a > b;
      cnt[S.slot_ab[k] + 1]++;
    }
  }
  for (int i = 0; i < nb; ++i) cnt[i + 1] += cnt[i];
  S.contrib_ptr = cnt;
  S.contrib.assign(cnt[nb], 0);
  std::vector<int> fill(cnt.begin(), cnt.end() - 1);
  for (int k = 0; k < E; ++k) {
    if (S.slot_aa[k] >= 0) S.contrib[fill[S.slot_aa[k]]++] = 4 * k + 0;
    if (S.slot_bb[k] >= 0) S.contrib[fill[S.slot_bb[k]]++] = 4 * k + 1;
    if (S.slot_ab[k] >= 0) S.contrib[fill[S.slot_ab[k]]++] = 4 * k + 2;
  }
}

// H blocks (nb x 9) and gradient (N x 3) at x, plus the total cost.
void assemble(const oracle_problem* p, const Structure& S, const double* x, double* Hv, double* g,
              double* cost_out, std::vector<double>& Jbuf, std::vector<double>& rbuf, SwitchBlocks* W = nullptr) {
  const int E = S.E, N = S.N;
  const int nb = (int)S.col_idx.size();
  Jbuf.resize((size_t)E * 18);
  rbuf.resize((size_t)E * 3);
  double cost = 0.0;
  const int nt = (W && W->on) ? 1 : set_threads(p);
  if (W && W->on) { W->hss.assign(E, 0.0); W->gs.assign(E, 0.0); W->hps.assign((size_t)E * 6, 0.0); }
  if (nt == 1) {
    // Serial, in residual-block order, like the reference's single-threaded Ceres.
    std::fill(Hv, Hv + (size_t)nb * 9, 0.0);
    std::fill(g, g + (size_t)N * 3, 0.0);
    for (int k = 0; k < E; ++k) {
      double* J = &Jbuf[(size_t)k * 18];
      double* r = &rbuf[(size_t)k * 3];
      double psi, rho1, c;
      if (W && W->has(p, k)) {
        double Js[3];
        const double sk = W->s[k];
        eval_edge_jet_sc(p, x, sk, k, r, J, Js, &c);
        const double pr = std::sqrt(W->lambda) * (1.0 - sk);          // prior residual, Jacobian -sqrt(lambda)
        c += 0.5 * pr * pr;
        W->hss[k] = Js[0] * Js[0] + Js[1] * Js[1] + Js[2] * Js[2] + W->lambda;
        W->gs[k] = Js[0] * r[0] + Js[1] * r[1] + Js[2] * r[2] - std::sqrt(W->lambda) * pr;
        for (int i = 0; i < 6; ++i) W->hps[(size_t)k * 6 + i] = J[i] * Js[0] + J[6 + i] * Js[1] + J[12 + i] * Js[2];
      } else {
        eval_edge_jet(p, x, k, false, r, J, &psi, &rho1, &c);
      }
      cost += c;
      const int a = p->edge_a[k], b = p->edge_b[k];
      if (S.slot_aa[k] >= 0) {
        double* H = Hv + (size_t)S.slot_aa[k] * 9;
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) H[3 * i + j] += J[i] * J[j] + J[6 + i] * J[6 + j] + J[12 + i] * J[12 + j];
          g[3 * a + i] += J[i] * r[0] + J[6 + i] * r[1] + J[12 + i] * r[2];
        }
      }
      if (S.slot_bb[k] >= 0) {
        double* H = Hv + (size_t)S.slot_bb[k] * 9;
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) H[3 * i + j] += J[3 + i] * J[3 + j] + J[9 + i] * J[9 + j] + J[15 + i] * J[15 + j];
          g[3 * b + i] += J[3 + i] * r[0] + J[9 + i] * r[1] + J[15 + i] * r[2];
        }
      }
      if (S.slot_ab[k] >= 0) {
        double* H = Hv + (size_t)S.slot_ab[k] * 9;
        const int oa = S.ab_transposed[k] ? 3 : 0, ob = S.ab_transposed[k] ? 0 : 3;
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j)
            H[3 * i + j] += J[oa + i] * J[ob + j] + J[6 + oa + i] * J[6 + ob + j] + J[12 + oa + i] * J[12 + ob + j];
      }
    }
    *cost_out = cost;
    return;
  }
  // Parallel: evaluate edges, then accumulate per block / per pose in edge order.
#pragma omp parallel for schedule(static) reduction(+ : cost)
  for (int k = 0; k < E; ++k) {
    double psi, rho1, c;
    eval_edge_jet(p, x, k, false, &rbuf[(size_t)k * 3], &Jbuf[(size_t)k * 18], &psi, &rho1, &c);
    cost += c;
  }
#pragma omp parallel for schedule(static)
  for (int s = 0; s < nb; ++s) {
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = S.contrib_ptr[s]; q < S.contrib_ptr[s + 1]; ++q) {
      const int k = S.contrib[q] >> 2, which = S.contrib[q] & 3;
      const double* J = &Jbuf[(size_t)k * 18];
      int oa, ob;
      if (which == 0) { oa = 0; ob = 0; } else if (which == 1) { oa = 3; ob = 3; }
      else { oa = S.ab_transposed[k] ? 3 : 0; ob = S.ab_transposed[k] ? 0 : 3; }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          H[3 * i + j] += J[oa + i] * J[ob + j] + J[6 + oa + i] * J[6 + ob + j] + J[12 + oa + i] * J[12 + ob + j];
    }
    std::memcpy(Hv + (size_t)s * 9, H, sizeof(H));
  }
  std::fill(g, g + (size_t)N * 3, 0.0);
  for (int k = 0; k < E; ++k) {  // cheap; serial keeps it deterministic
    const double* J = &Jbuf[(size_t)k * 18];
    const double* r = &rbuf[(size_t)k * 3];
    const int a = p->edge_a[k], b = p->edge_b[k];
    if (S.slot_aa[k] >= 0) for (int i = 0; i < 3; ++i) g[3 * a + i] += J[i] * r[0] + J[6 + i] * r[1] + J[12 + i] * r[2];
    if (S.slot_bb[k] >= 0) for (int i = 0; i < 3; ++i) g[3 * b + i] += J[3 + i] * r[0] + J[9 + i] * r[1] + J[15 + i] * r[2];
  }
  *cost_out = cost;
}

double total_cost(const oracle_problem* p, const double* x) {
  const int E = p->n_edges;
  double cost = 0.0;
  const int nt = set_threads(p);
  if (nt == 1) { for (int k = 0; k < E; ++k) cost += cost_edge_double(p, x, k); return cost; }
#pragma omp parallel for schedule(static) reduction(+ : cost)
  for (int k = 0; k < E; ++k) cost += cost_edge_double(p, x, k);
  return cost;
}

// ------------------------------------------------------------------------------------
// Exact sparse Cholesky of the scalar normal matrix (own implementation of the textbook
// elimination-tree / up-looking algorithm, minimum-degree ordering on the pose graph).
// Stands in for Ceres' SPARSE_NORMAL_CHOLESKY backend (CHOLMOD / Eigen SimplicialLDLT).
// ------------------------------------------------------------------------------------
struct SparseChol {
  int n = 0;                       // scalar dimension = 3 * #free poses
  std::vector<int> pose_to_ord;    // pose -> position in elimination order (-1 if not free)
  std::vector<int> ord_to_pose;
  // permuted upper-triangular CSC of A
  std::vector<int> Ap, Ai;
  std::vector<double> Ax;
  // mapping H block value (slot*9 + r*3 + c) -> Ax position (-1 if lower part of a diag block)
  std::vector<int> map_H;
  std::vector<int> diag_pos;       // Ax position of scalar diagonal i
  std::vector<int> parent;
  // L by columns (rows > col), diagonal separate
  std::vector<std::vector<int>> Li;
  std::vector<std::vector<double>> Lx;
  std::vector<double> Ld;
  int64_t nnzL = 0;

  void order(const Structure& S) {
    const int N = S.N;
    std::vector<std::unordered_set<int>> adj(N);
    for (int i = 0; i < N; ++i)
      for (int q = S.row_ptr[i]; q < S.row_ptr[i + 1]; ++q) {
        const int j = S.col_idx[q];
        if (j != i) { adj[i].insert(j); adj[j].insert(i); }
      }
    typedef std::pair<int, int> DI;
    std::priority_queue<DI, std::vector<DI>, std::greater<DI>> pq;
    std::vector<char> done(N, 0);
    int nfree = 0;
    for (int i = 0; i < N; ++i) if (S.is_free[i]) { pq.push(DI((int)adj[i].size(), i)); ++nfree; } else done[i] = 1;
    pose_to_ord.assign(N, -1);
    ord_to_pose.clear();
    ord_to_pose.reserve(nfree);
    while (!pq.empty()) {
      const DI top = pq.top(); pq.pop();
      const int v = top.second;
      if (done[v] || top.first != (int)adj[v].size()) continue;
      done[v] = 1;
      pose_to_ord[v] = (int)ord_to_pose.size();
      ord_to_pose.push_back(v);
      std::vector<int> nb(adj[v].begin(), adj[v].end());
      for (int u : nb) adj[u].erase(v);
      for (size_t i = 0; i < nb.size(); ++i)
        for (size_t j = i + 1; j < nb.size(); ++j) { adj[nb[i]].insert(nb[j]); adj[nb[j]].insert(nb[i]); }
      for (int u : nb) pq.push(DI((int)adj[u].size(), u));
      std::unordered_set<int>().swap(adj[v]);
    }
    n = 3 * nfree;
  }

  void symbolic(const Structure& S) {
    order(S);
    const int nb = (int)S.col_idx.size();
    // count entries per permuted column
    std::vector<int> cnt(n + 1, 0);
    struct Ent { int row, col, src; };
    std::vector<Ent> ents;
    ents.reserve((size_t)nb * 9);
    for (int i = 0; i < S.N; ++i)
      for (int q = S.row_ptr[i]; q < S.row_ptr[i + 1]; ++q) {
        const int j = S.col_idx[q];
        const int oi = pose_to_ord[i], oj = pose_to_ord[j];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            if (i == j && r > c) continue;
            int pr = 3 * oi + r, pc = 3 * oj + c;
            if (pr > pc) std::swap(pr, pc);
            ents.push_back({pr, pc, q * 9 + r * 3 + c});
          }
      }
    std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.col != b.col ? a.col < b.col : a.row < b.row; });
    Ap.assign(n + 1, 0);
    Ai.resize(ents.size());
    Ax.assign(ents.size(), 0.0);
    map_H.assign((size_t)nb * 9, -1);
    diag_pos.assign(n, -1);
    for (size_t t = 0; t < ents.size(); ++t) {
      Ap[ents[t].col + 1]++;
      Ai[t] = ents[t].row;
      map_H[ents[t].src] = (int)t;
      if (ents[t].row == ents[t].col) diag_pos[ents[t].col] = (int)t;
    }
    for (int c = 0; c < n; ++c) Ap[c + 1] += Ap[c];
    // elimination tree
    parent.assign(n, -1);
    std::vector<int> anc(n, -1);
    for (int k = 0; k < n; ++k)
      for (int q = Ap[k]; q < Ap[k + 1]; ++q) {
        int i = Ai[q];
        while (i != -1 && i < k) {
          const int nx = anc[i];
          anc[i] = k;
          if (nx == -1) parent[i] = k;
          i = nx;
        }
      }
    Li.assign(n, std::vector<int>());
    Lx.assign(n, std::vector<double>());
    Ld.assign(n, 0.0);
  }

  void load(const double* Hv, const double* diag_add /* n-vector in permuted scalar order */) {
    std::fill(Ax.begin(), Ax.end(), 0.0);
    for (size_t t = 0; t < map_H.size(); ++t) if (map_H[t] >= 0) Ax[map_H[t]] += Hv[t];
    if (diag_add) for (int i = 0; i < n; ++i) Ax[diag_pos[i]] += diag_add[i];
  }

  bool factor() {
    std::vector<double> x(n, 0.0);
    std::vector<int> mark(n, -1), stack(n), path(n);
    for (int k = 0; k < n; ++k) { Li[k].clear(); Lx[k].clear(); }
    nnzL = 0;
    for (int k = 0; k < n; ++k) {
      int top = n;
      mark[k] = k;
      double d = 0.0;
      for (int q = Ap[k]; q < Ap[k + 1]; ++q) {
        int i = Ai[q];
        if (i == k) { d = Ax[q]; continue; }
        x[i] = Ax[q];
        int len = 0;
        while (mark[i] != k) { path[len++] = i; mark[i] = k; i = parent[i]; }
        while (len > 0) stack[--top] = path[--len];
      }
      for (; top < n; ++top) {
        const int j = stack[top];
        const double lkj = x[j] / Ld[j];
        x[j] = 0.0;
        const std::vector<int>& li = Li[j];
        const std::vector<double>& lx = Lx[j];
        for (size_t q = 0; q < li.size(); ++q) x[li[q]] -= lx[q] * lkj;
        d -= lkj * lkj;
        Li[j].push_back(k);
        Lx[j].push_back(lkj);
        ++nnzL;
      }
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      Ld[k] = std::sqrt(d);
    }
    nnzL += n;
    return true;
  }

  void solve(double* b) const {  // in place, permuted scalar order
    for (int j = 0; j < n; ++j) {
      b[j] /= Ld[j];
      const double bj = b[j];
      for (size_t q = 0; q < Li[j].size(); ++q) b[Li[j][q]] -= Lx[j][q] * bj;
    }
    for (int j = n - 1; j >= 0; --j) {
      double s = b[j];
      for (size_t q = 0; q < Li[j].size(); ++q) s -= Lx[j][q] * b[Li[j][q]];
      b[j] = s / Ld[j];
    }
  }
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool check_problem(const oracle_problem* p) {
  if (!p || p->n_poses <= 0 || p->n_edges < 0 || !p->pose_xyt) return false;
  if (p->n_edges > 0 && (!p->edge_a || !p->edge_b || !p->meas_xyt || !p->kind)) return false;
  for (int k = 0; k < p->n_edges; ++k) {
    const int a = p->edge_a[k], b = p->edge_b[k];
    if (a < 0 || b < 0 || a >= p->n_poses || b >= p->n_poses || a == b) return false;
  }
  return true;
}

}  // namespace

extern "C" {

void oracle_lm_options_default(oracle_lm_options* o) {
  o->max_num_iterations = 50;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->verbose = 0;
}

int oracle_evaluate(const oracle_problem* p, const double* pose_xyt, int raw, double* cost,
                    double* residuals, double* jacobians, double* psi, double* rho1, double* gradient) {
  if (!check_problem(p)) return 1;
  const double* x = pose_xyt ? pose_xyt : p->pose_xyt;
  const int E = p->n_edges, N = p->n_poses;
  if (gradient) std::fill(gradient, gradient + (size_t)N * 3, 0.0);
  std::vector<char> is_free(N, 0);
  for (int k = 0; k < E; ++k) { is_free[p->edge_a[k]] = 1; is_free[p->edge_b[k]] = 1; }
  if (p->fixed_pose >= 0 && p->fixed_pose < N) is_free[p->fixed_pose] = 0;
  double total = 0.0;
  for (int k = 0; k < E; ++k) {
    double r[3], J[18], ps, r1, c;
    eval_edge_jet(p, x, k, raw != 0, r, J, &ps, &r1, &c);
    total += c;
    if (residuals) std::memcpy(residuals + (size_t)3 * k, r, sizeof(r));
    if (jacobians) std::memcpy(jacobians + (size_t)18 * k, J, sizeof(J));
    if (psi) psi[k] = ps;
    if (rho1) rho1[k] = r1;
    if (gradient) {
      const int a = p->edge_a[k], b = p->edge_b[k];
      if (is_free[a]) for (int i = 0; i < 3; ++i) gradient[3 * a + i] += J[i] * r[0] + J[6 + i] * r[1] + J[12 + i] * r[2];
      if (is_free[b]) for (int i = 0; i < 3; ++i) gradient[3 * b + i] += J[3 + i] * r[0] + J[9 + i] * r[1] + J[15 + i] * r[2];
    }
  }
  if (cost) *cost = total;
  return 0;
}

int oracle_cost(const oracle_problem* p, const double* pose_xyt, double* cost) {
  if (!check_problem(p) || !cost) return 1;
  *cost = total_cost(p, pose_xyt ? pose_xyt : p->pose_xyt);
  return 0;
}

// Independent closed form: u = Ra^T (tb - ta) - tm, (ex,ey) = Rm^T u,
// e_th = asin(sin(thb - tha - thm)); analytic Jacobian; DCS rank-1 term.
void oracle_edge_closed_form(const double* pa, const double* pb, const double* meas, int dcs,
                             double phi, double* e, double* J, double* psi_out) {
  const double ca = std::cos(pa[2]), sa = std::sin(pa[2]);
  const double cm = std::cos(meas[2]), sm = std::sin(meas[2]);
  const double dxw = pb[0] - pa[0], dyw = pb[1] - pa[1];
  const double px = ca * dxw + sa * dyw, py = -sa * dxw + ca * dyw;   // Ra^T d
  const double ux = px - meas[0], uy = py - meas[1];
  const double ex = cm * ux + sm * uy, ey = -sm * ux + cm * uy;
  const double delta = pb[2] - pa[2] - meas[2];
  const double sd = std::sin(delta), cd = std::cos(delta);
  const double eth = std::asin(sd);
  const double sig = cd / std::sqrt(1.0 - sd * sd);
  // Q = Rm^T Ra^T
  const double q00 = cm * ca - sm * sa, q01 = cm * sa + sm * ca;   // cos(tha+thm), sin(tha+thm)
  const double q10 = -q01, q11 = q00;
  // d(Ra^T d)/dtha = (py, -px); w = Rm^T (py, -px)
  const double wx = cm * py - sm * px, wy = -sm * py - cm * px;
  double Jp[18] = {-q00, -q01, wx, q00, q01, 0.0,
                   -q10, -q11, wy, q10, q11, 0.0,
                   0.0, 0.0, -sig, 0.0, 0.0, sig};
  double r[3] = {ex, ey, eth};
  double psi = 1.0;
  if (dcs) {
    const double res = ex * ex + ey * ey;
    const double psi_org = std::sqrt(2.0 * phi / (phi + res));
    if (psi_org < 1.0) {
      psi = psi_org;
      const double kap = -psi_org / (phi + res);
      double gp[6];
      for (int j = 0; j < 6; ++j) gp[j] = kap * (ex * Jp[j] + ey * Jp[6 + j]);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) Jp[6 * i + j] = psi * Jp[6 * i + j] + r[i] * gp[j];
      for (int i = 0; i < 3; ++i) r[i] *= psi;
    }
  }
  for (int i = 0; i < 3; ++i) e[i] = r[i];
  for (int i = 0; i < 18; ++i) J[i] = Jp[i];
  if (psi_out) *psi_out = psi;
}

// SwitchableClosureResidue (ceres_error.cpp:238-297): the plain functor's diff, then e = s * (...) with s the
// seventh Jet parameter.  Jet product rule h.v = f.a g.v + f.v g.a with f = s: pose partials s * d.v + 0 * d.a,
// switch partial s * 0 + 1 * d.a - written out so the floating-point operations are the reference's.
void oracle_sc_edge(const double* pa, const double* pb, const double* meas, double s, double* e, double* J) {
  const Meas q = make_meas(meas[0], meas[1], meas[2]);
  Jet P1[3], P2[3], d[3], ps;
  for (int i = 0; i < 3; ++i) { P1[i] = Jet(pa[i], i); P2[i] = Jet(pb[i], 3 + i); }
  functor<Jet>(q, false, 0.5, P1, P2, d, &ps);
  for (int i = 0; i < 3; ++i) {
    e[i] = s * d[i].a;
    if (J) {
      for (int j = 0; j < 6; ++j) J[7 * i + j] = s * d[i].v[j] + 0.0 * d[i].a;
      J[7 * i + 6] = s * 0.0 + 1.0 * d[i].a;
    }
  }
}
// SwitchPriorResidue (:300-317): T(sqrt(lambda)) * (T(1) - S[0])
void oracle_sc_prior(double lambda, double s, double* e, double* J) {
  const double w = std::sqrt(lambda);
  e[0] = w * (1.0 - s);
  if (J) J[0] = w * (0.0 - 1.0) + 0.0 * (1.0 - s);
}

int oracle_pattern(const oracle_problem* p, int32_t* nnzb, int32_t* row_ptr, int32_t* col_idx) {
  if (!check_problem(p)) return 1;
  Structure S;
  build_structure(p, S);
  if (nnzb) *nnzb = (int32_t)S.col_idx.size();
  if (row_ptr) std::copy(S.row_ptr.begin(), S.row_ptr.end(), row_ptr);
  if (col_idx) std::copy(S.col_idx.begin(), S.col_idx.end(), col_idx);
  return 0;
}

int oracle_hessian(const oracle_problem* p, const double* pose_xyt, double* block_values, double* gradient) {
  if (!check_problem(p)) return 1;
  Structure S;
  build_structure(p, S);
  std::vector<double> Hv(S.col_idx.size() * 9), g((size_t)p->n_poses * 3), Jb, rb;
  double cost;
  assemble(p, S, pose_xyt ? pose_xyt : p->pose_xyt, Hv.data(), g.data(), &cost, Jb, rb);
  if (block_values) std::copy(Hv.begin(), Hv.end(), block_values);
  if (gradient) std::copy(g.begin(), g.end(), gradient);
  return 0;
}

double oracle_time_linearize(const oracle_problem* p, int repeats) {
  if (!check_problem(p) || repeats <= 0) return -1.0;
  Structure S;
  build_structure(p, S);
  std::vector<double> Hv(S.col_idx.size() * 9), g((size_t)p->n_poses * 3), Jb, rb;
  double cost;
  assemble(p, S, p->pose_xyt, Hv.data(), g.data(), &cost, Jb, rb);  // warm-up
  const double t0 = now_s();
  for (int i = 0; i < repeats; ++i) assemble(p, S, p->pose_xyt, Hv.data(), g.data(), &cost, Jb, rb);
  return (now_s() - t0) / repeats;
}

int oracle_linear_solve(const oracle_problem* p, const double* pose_xyt, const double* lambda,
                        const double* rhs, double* w) {
  if (!check_problem(p) || !rhs || !w) return 1;
  Structure S;
  build_structure(p, S);
  std::vector<double> Hv(S.col_idx.size() * 9), g((size_t)p->n_poses * 3), Jb, rb;
  double cost;
  assemble(p, S, pose_xyt ? pose_xyt : p->pose_xyt, Hv.data(), g.data(), &cost, Jb, rb);
  SparseChol C;
  C.symbolic(S);
  std::vector<double> dadd(C.n, 0.0), b(C.n, 0.0);
  for (int o = 0; o < C.n / 3; ++o) {
    const int i = C.ord_to_pose[o];
    for (int c = 0; c < 3; ++c) { dadd[3 * o + c] = lambda ? lambda[3 * i + c] : 0.0; b[3 * o + c] = rhs[3 * i + c]; }
  }
  C.load(Hv.data(), dadd.data());
  if (!C.factor()) return 2;
  C.solve(b.data());
  std::fill(w, w + (size_t)p->n_poses * 3, 0.0);
  for (int o = 0; o < C.n / 3; ++o) {
    const int i = C.ord_to_pose[o];
    for (int c = 0; c < 3; ++c) w[3 * i + c] = b[3 * o + c];
  }
  return 0;
}

static int lm_impl(const oracle_problem* p, const oracle_lm_options* opt, double* pose_xyt_inout, double* switch_inout,
                   double sc_lambda, oracle_summary* sum, oracle_iteration* trace, int32_t trace_cap) {
  if (!check_problem(p) || !opt || !pose_xyt_inout || !sum) return 1;
  std::memset(sum, 0, sizeof(*sum));
  const int N = p->n_poses;
  const double t_start = now_s();
  Structure S;
  build_structure(p, S);
  SparseChol C;
  C.symbolic(S);
  const int nb = (int)S.col_idx.size();
  const int n = C.n;
  const int E = p->n_edges;
  // METHOD 2: one switch per loop edge, eliminated edge by edge inside the linear solve (a switch couples only
  // to the two poses of its edge, so the Schur complement onto the poses keeps the pose block pattern and is
  // what a sparse Cholesky of the full system computes)
  SwitchBlocks W;
  W.on = switch_inout != nullptr;
  W.lambda = sc_lambda;
  if (W.on) W.s.assign(switch_inout, switch_inout + E);
  std::vector<double> sw_c(W.s), sw_best(W.s), sw_scale(W.on ? E : 0, 1.0), sw_diag(W.on ? E : 0, 0.0), sw_step(W.on ? E : 0, 0.0);
  std::vector<double> Hr;

  std::vector<double> x(pose_xyt_inout, pose_xyt_inout + (size_t)N * 3), xc(x), best(x);
  std::vector<double> Hv((size_t)nb * 9), g((size_t)N * 3), Jb, rb;
  std::vector<double> scale((size_t)N * 3, 1.0), diagonal((size_t)N * 3, 0.0);
  std::vector<double> Hs((size_t)nb * 9), dadd(n), b(n), step((size_t)N * 3, 0.0), delta((size_t)N * 3, 0.0);
  double eval_time = 0.0, solve_time = 0.0;

  auto free_norm = [&](const std::vector<double>& v, const std::vector<double>& sv) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) if (S.is_free[i]) for (int c = 0; c < 3; ++c) s += v[3 * i + c] * v[3 * i + c];
    for (int k = 0; k < E; ++k) if (W.has(p, k)) s += sv[k] * sv[k];
    return std::sqrt(s);
  };
  auto grad_norms = [&](double* gmax, double* gnorm) {
    double m = 0.0, s = 0.0;
    for (int i = 0; i < N; ++i) if (S.is_free[i]) for (int c = 0; c < 3; ++c) { m = std::max(m, std::fabs(g[3 * i + c])); s += g[3 * i + c] * g[3 * i + c]; }
    for (int k = 0; k < E; ++k) if (W.has(p, k)) { m = std::max(m, std::fabs(W.gs[k])); s += W.gs[k] * W.gs[k]; }
    *gmax = m; *gnorm = std::sqrt(s);
  };
  auto cost_at = [&](const double* xx, const std::vector<double>& sv) {
    if (!W.on) return total_cost(p, xx);
    double c = 0.0;
    for (int k = 0; k < E; ++k) c += W.has(p, k) ? cost_edge_double_sc(p, xx, sv[k], W.lambda, k) : cost_edge_double(p, xx, k);
    return c;
  };
  auto hdiag = [&](int i, int c) { return Hv[(size_t)S.row_ptr[i] * 9 + 4 * c]; };  // diag block is first in its row

  int n_logged = 0;
  double min_logged_cost = std::numeric_limits<double>::max();
  auto log_iter = [&](const oracle_iteration& it) {
    if (trace && n_logged < trace_cap) trace[n_logged] = it;
    ++n_logged;
    min_logged_cost = std::min(min_logged_cost, it.cost);
    if (opt->verbose) {
      if (it.iteration == 0)
        std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time\n");
      std::printf("% 4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n", it.iteration, it.cost,
                  it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius, 1,
                  it.iteration_time_s, it.cumulative_time_s);
    }
  };

  // --- iteration zero -----------------------------------------------------------------
  double x_cost = 0.0;
  double x_norm = free_norm(x, W.s);
  double t0 = now_s();
  assemble(p, S, x.data(), Hv.data(), g.data(), &x_cost, Jb, rb, &W);
  eval_time += now_s() - t0;
  if (!std::isfinite(x_cost)) { sum->termination_type = 2; std::snprintf(sum->message, sizeof(sum->message), "Initial cost is not finite."); return 4; }
  if (opt->jacobi_scaling)
  {
    for (int i = 0; i < N; ++i) if (S.is_free[i]) for (int c = 0; c < 3; ++c) scale[3 * i + c] = 1.0 / (1.0 + std::sqrt(hdiag(i, c)));
    for (int k = 0; k < E; ++k) if (W.has(p, k)) sw_scale[k] = 1.0 / (1.0 + std::sqrt(W.hss[k]));
  }
  oracle_iteration it;
  std::memset(&it, 0, sizeof(it));
  it.iteration = 0;
  it.cost = x_cost;
  grad_norms(&it.gradient_max_norm, &it.gradient_norm);
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  it.trust_region_radius = radius;
  it.iteration_time_s = now_s() - t_start;
  it.cumulative_time_s = it.iteration_time_s;
  sum->initial_cost = x_cost;
  double minimum_cost = x_cost;
  int invalid = 0;
  int term = 1;  // NO_CONVERGENCE
  const char* msg = "Maximum number of iterations reached.";
  log_iter(it);
  oracle_iteration prev = it;

  if (it.gradient_max_norm <= opt->gradient_tolerance) { term = 0; msg = "Gradient tolerance reached."; }
  else
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (checks on the last logged iteration)
    if (prev.iteration >= opt->max_num_iterations) { term = 1; msg = "Maximum number of iterations reached."; break; }
    if (prev.gradient_max_norm <= opt->gradient_tolerance) { term = 0; msg = "Gradient tolerance reached."; break; }
    if (radius <= opt->min_trust_region_radius) { term = 0; msg = "Minimum trust region radius reached."; break; }

    const double it_start = now_s();
    std::memset(&it, 0, sizeof(it));
    it.iteration = prev.iteration + 1;

    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal)
      for (int i = 0; i < N; ++i) if (S.is_free[i]) for (int c = 0; c < 3; ++c) {
        const double d = scale[3 * i + c] * scale[3 * i + c] * hdiag(i, c);   // colnorm^2 of the scaled J
        diagonal[3 * i + c] = std::min(std::max(d, opt->min_lm_diagonal), opt->max_lm_diagonal);
      }
    if (!reuse_diagonal)
      for (int k = 0; k < E; ++k) if (W.has(p, k))
        sw_diag[k] = std::min(std::max(sw_scale[k] * sw_scale[k] * W.hss[k], opt->min_lm_diagonal), opt->max_lm_diagonal);
    t0 = now_s();
    // scaled normal equations: Hs = S H S, rhs = S g, D^2 = diagonal / radius
    for (int i = 0; i < N; ++i)
      for (int q = S.row_ptr[i]; q < S.row_ptr[i + 1]; ++q) {
        const int j = S.col_idx[q];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
          Hs[(size_t)q * 9 + 3 * r + c] = Hv[(size_t)q * 9 + 3 * r + c] * scale[3 * i + r] * scale[3 * j + c];
      }
    for (int o = 0; o < n / 3; ++o) {
      const int i = C.ord_to_pose[o];
      for (int c = 0; c < 3; ++c) {
        const double lm = std::sqrt(diagonal[3 * i + c] / radius);
        dadd[3 * o + c] = lm * lm;
        b[3 * o + c] = scale[3 * i + c] * g[3 * i + c];
      }
    }
    // METHOD 2: Schur complement of every switch onto its two poses (scaled variables)
    const double* Hload = Hs.data();
    std::vector<double> sw_den, sw_b, sw_u;
    if (W.on) {
      Hr = Hs;
      sw_den.assign(E, 1.0); sw_b.assign(E, 0.0); sw_u.assign((size_t)E * 6, 0.0);
      for (int k = 0; k < E; ++k) if (W.has(p, k)) {
        const int ia = p->edge_a[k], ib = p->edge_b[k];
        const double sg_ = sw_scale[k];
        const double den = sg_ * sg_ * W.hss[k] + sw_diag[k] / radius;
        double* u = &sw_u[(size_t)k * 6];
        for (int c = 0; c < 3; ++c) { u[c] = W.hps[(size_t)k * 6 + c] * scale[3 * ia + c] * sg_; u[3 + c] = W.hps[(size_t)k * 6 + 3 + c] * scale[3 * ib + c] * sg_; }
        const double bs = sg_ * W.gs[k];
        sw_den[k] = den; sw_b[k] = bs;
        if (S.slot_aa[k] >= 0) {
          double* H = &Hr[(size_t)S.slot_aa[k] * 9];
          const int o = C.pose_to_ord[ia];
          for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) H[3 * r + c] -= u[r] * u[c] / den; b[3 * o + r] -= u[r] * bs / den; }
        }
        if (S.slot_bb[k] >= 0) {
          double* H = &Hr[(size_t)S.slot_bb[k] * 9];
          const int o = C.pose_to_ord[ib];
          for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) H[3 * r + c] -= u[3 + r] * u[3 + c] / den; b[3 * o + r] -= u[3 + r] * bs / den; }
        }
        if (S.slot_ab[k] >= 0) {
          double* H = &Hr[(size_t)S.slot_ab[k] * 9];
          const int oa = S.ab_transposed[k] ? 3 : 0, ob = S.ab_transposed[k] ? 0 : 3;
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] -= u[oa + r] * u[ob + c] / den;
        }
      }
      Hload = Hr.data();
    }
    C.load(Hload, dadd.data());
    bool ok = C.factor();
    if (ok) C.solve(b.data());
    reuse_diagonal = true;
    solve_time += now_s() - t0;
    sum->factor_nnz = C.nnzL;
    if (ok) for (int i = 0; i < n; ++i) if (!std::isfinite(b[i])) { ok = false; break; }
    double model_cost_change = 0.0;
    if (ok) {
      std::fill(step.begin(), step.end(), 0.0);
      for (int o = 0; o < n / 3; ++o) { const int i = C.ord_to_pose[o]; for (int c = 0; c < 3; ++c) step[3 * i + c] = -b[3 * o + c]; }
      // back-substitution of the switches: y_s = (b_s - u^T y_p) / den, step = -y
      for (int k = 0; k < E; ++k) if (W.has(p, k)) {
        const int ia = p->edge_a[k], ib = p->edge_b[k];
        const double* u = &sw_u[(size_t)k * 6];
        double uy = 0.0;   // u^T y_p = -u^T step_p   (constant poses have step 0)
        for (int c = 0; c < 3; ++c) uy -= u[c] * step[3 * ia + c] + u[3 + c] * step[3 * ib + c];
        sw_step[k] = -(sw_b[k] - uy) / sw_den[k];
      }
      // model_cost_change = -(J step)^T (r + J step / 2) = -step^T (S g) - step^T Hs step / 2
      double sg = 0.0, shs = 0.0;
      for (int i = 0; i < N; ++i) if (S.is_free[i]) for (int c = 0; c < 3; ++c) sg += step[3 * i + c] * scale[3 * i + c] * g[3 * i + c];
      for (int i = 0; i < N; ++i)
        for (int q = S.row_ptr[i]; q < S.row_ptr[i + 1]; ++q) {
          const int j = S.col_idx[q];
          double acc = 0.0;
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) acc += step[3 * i + r] * Hs[(size_t)q * 9 + 3 * r + c] * step[3 * j + c];
          shs += (i == j) ? acc : 2.0 * acc;
        }
      for (int k = 0; k < E; ++k) if (W.has(p, k)) {
        const int ia = p->edge_a[k], ib = p->edge_b[k];
        const double* u = &sw_u[(size_t)k * 6];
        double us = 0.0;
        for (int c = 0; c < 3; ++c) us += u[c] * step[3 * ia + c] + u[3 + c] * step[3 * ib + c];
        sg += sw_step[k] * sw_b[k];
        shs += sw_scale[k] * sw_scale[k] * W.hss[k] * sw_step[k] * sw_step[k] + 2.0 * sw_step[k] * us;
      }
      model_cost_change = -sg - 0.5 * shs;
      it.step_is_valid = model_cost_change > 0.0;
    }
    if (!it.step_is_valid) {
      // HandleInvalidStep
      if (++invalid >= opt->max_num_consecutive_invalid_steps) { term = 2; msg = "Number of consecutive invalid steps more than max_num_consecutive_invalid_steps."; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;   // StepRejected(0)
      it.cost = x_cost; it.cost_change = 0.0;
      it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;
      it.step_norm = 0.0; it.relative_decrease = 0.0;
      it.trust_region_radius = radius;
      it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
      log_iter(it); prev = it; sum->num_unsuccessful_steps++;
      continue;
    }
    invalid = 0;
    for (size_t i = 0; i < delta.size(); ++i) delta[i] = step[i] * scale[i];
    for (size_t i = 0; i < xc.size(); ++i) xc[i] = x[i] + delta[i];
    for (int k = 0; k < E; ++k) if (W.has(p, k)) sw_c[k] = W.s[k] + sw_step[k] * sw_scale[k];
    t0 = now_s();
    double cand = cost_at(xc.data(), sw_c);
    eval_time += now_s() - t0;
    if (!std::isfinite(cand)) cand = std::numeric_limits<double>::max();

    // ParameterToleranceReached
    { double s = 0.0; for (int i = 0; i < N; ++i) if (S.is_free[i]) for (int c = 0; c < 3; ++c) { const double d = x[3 * i + c] - xc[3 * i + c]; s += d * d; }
      for (int k = 0; k < E; ++k) if (W.has(p, k)) { const double d = W.s[k] - sw_c[k]; s += d * d; }
      it.step_norm = std::sqrt(s); }
    it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;
    const double step_size_tolerance = opt->parameter_tolerance * (x_norm + opt->parameter_tolerance);
    if (it.step_norm <= step_size_tolerance) {
      term = 0; msg = "Parameter tolerance reached.";
      it.cost = x_cost; it.trust_region_radius = radius;
      it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
      break;   // Ceres returns from Minimize() here: the terminating iteration is not appended to summary.iterations
    }
    // FunctionToleranceReached
    it.cost_change = x_cost - cand;
    if (std::fabs(it.cost_change) <= opt->function_tolerance * x_cost) {
      term = 0; msg = "Function tolerance reached.";
      it.cost = x_cost; it.trust_region_radius = radius;
      it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
      break;   // Ceres returns from Minimize() here: the terminating iteration is not appended to summary.iterations
    }
    // IsStepSuccessful
    it.relative_decrease = (cand >= std::numeric_limits<double>::max()) ? std::numeric_limits<double>::lowest()
                                                                          : (x_cost - cand) / model_cost_change;
    if (it.relative_decrease > opt->min_relative_decrease) {
      // HandleSuccessfulStep
      x = xc;
      if (W.on) W.s = sw_c;
      x_norm = free_norm(x, W.s);
      t0 = now_s();
      assemble(p, S, x.data(), Hv.data(), g.data(), &x_cost, Jb, rb, &W);
      eval_time += now_s() - t0;
      it.step_is_successful = 1;
      it.cost = x_cost;
      grad_norms(&it.gradient_max_norm, &it.gradient_norm);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt->max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      sum->num_successful_steps++;
      if (x_cost < minimum_cost) { minimum_cost = x_cost; best = x; sw_best = W.s; }
    } else {
      it.step_is_successful = 0;
      it.cost = cand;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      sum->num_unsuccessful_steps++;
    }
    it.trust_region_radius = radius;
    it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
    log_iter(it);
    prev = it;
  }

  // Solver::Summary: user parameters hold the best accepted iterate; final_cost is the
  // minimum logged cost (Ceres SetSummaryFinalCost).
  std::copy(best.begin(), best.end(), pose_xyt_inout);
  if (W.on) std::copy(sw_best.begin(), sw_best.end(), switch_inout);
  sum->final_cost = std::min(sum->initial_cost, min_logged_cost);
  sum->num_iterations = n_logged;
  sum->termination_type = term;
  sum->total_time_s = now_s() - t_start;
  sum->eval_time_s = eval_time;
  sum->linear_solver_time_s = solve_time;
  std::snprintf(sum->message, sizeof(sum->message), "%s", msg);
  return 0;
}


int oracle_solve(const oracle_problem* p, const oracle_lm_options* opt, double* pose_xyt_inout,
                 oracle_summary* sum, oracle_iteration* trace, int32_t trace_cap) {
  return lm_impl(p, opt, pose_xyt_inout, nullptr, 1.0, sum, trace, trace_cap);
}

// METHOD 2 (main.cpp:105-150 with SC_ON): switches[E] in/out (entries of odometry edges are ignored; the
// reference starts every switch at 1), prior weight lambda (1.0 in the reference).  p->dcs_on must be 0.
int oracle_sc_solve(const oracle_problem* p, double lambda, const oracle_lm_options* opt, double* pose_xyt_inout,
                    double* switches_inout, oracle_summary* sum, oracle_iteration* trace, int32_t trace_cap) {
  if (!p || p->dcs_on || !switches_inout) return 1;
  return lm_impl(p, opt, pose_xyt_inout, switches_inout, lambda, sum, trace, trace_cap);
}

}  // extern "C"

// C wrapper around the REFERENCE's own residual functors (DCS-ceres/include/ceres_error.h,
// DCS-ceres/src/ceres_error.cpp), compiled from /root/reference where they lie against the
// Eigen/Ceres stand-ins in this directory.  Output goes to oracle/_ref/ (git-ignored).
// TEST INFRASTRUCTURE ONLY: used to pin oracle/dcs_oracle.cpp and to generate tests/golden/.
#include <utility>
#include "ceres_error.h"

namespace {
int eval33(ceres::CostFunction* cf, const double* pa, const double* pb, double* e, double* J) {
  const double* params[2] = {pa, pb};
  double Ja[9], Jb[9];
  double* jac[2] = {Ja, Jb};
  const bool ok = cf->Evaluate(params, e, J ? jac : nullptr);
  if (ok && J)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) { J[6 * r + c] = Ja[3 * r + c]; J[6 * r + 3 + c] = Jb[3 * r + c]; }
  delete cf;
  return ok ? 0 : 1;
}
}  // namespace

extern "C" {
// kind: 0 = OdometryResidue, 1 = DCSClosureResidue. J: 3x6 row-major or NULL (double path).
int ref_edge(int kind, const double* meas, const double* pa, const double* pb, double* e, double* J) {
  ceres::CostFunction* cf = kind ? DCSClosureResidue::Create(meas[0], meas[1], meas[2])
                                 : OdometryResidue::Create(meas[0], meas[1], meas[2]);
  return eval33(cf, pa, pb, e, J);
}
int ref_edges(int n, const unsigned char* kind, const double* meas, const double* pa, const double* pb,
              double* e, double* J) {
  for (int k = 0; k < n; ++k)
    if (ref_edge(kind[k], meas + 3 * k, pa + 3 * k, pb + 3 * k, e + 3 * k, J ? J + 18 * k : nullptr)) return 1;
  return 0;
}
// SwitchableClosureResidue (METHOD 2): J 3x7 row-major (pa, pb, s).
int ref_switchable_edge(const double* meas, const double* pa, const double* pb, double s, double* e, double* J) {
  ceres::CostFunction* cf = SwitchableClosureResidue::Create(meas[0], meas[1], meas[2]);
  const double* params[3] = {pa, pb, &s};
  double Ja[9], Jb[9], Js[3];
  double* jac[3] = {Ja, Jb, Js};
  const bool ok = cf->Evaluate(params, e, J ? jac : nullptr);
  if (ok && J)
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) { J[7 * r + c] = Ja[3 * r + c]; J[7 * r + 3 + c] = Jb[3 * r + c]; }
      J[7 * r + 6] = Js[r];
    }
  delete cf;
  return ok ? 0 : 1;
}
int ref_switch_prior(double lambda, double s, double* e, double* J) {
  ceres::CostFunction* cf = SwitchPriorResidue::Create(lambda);
  const double* params[1] = {&s};
  double* jac[1] = {J};
  const bool ok = cf->Evaluate(params, e, J ? jac : nullptr);
  delete cf;
  return ok ? 0 : 1;
}
}

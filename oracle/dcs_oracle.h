/*
 * dcs_oracle.h — CPU restatement of the reference's DCS-ceres solve path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: the reference's functor text (src/ceres_error.cpp) is pinned — the
 * restatement in dcs_oracle.cpp is checked against the reference's own functor sources
 * compiled (where they lie under /root/reference) against the Jet/3x3-matrix shims in
 * oracle/ref_shim/ -> oracle/_ref/libdcs_ref.so, and against closed forms / finite
 * differences.  The minimiser (Ceres trust-region LM + SPARSE_NORMAL_CHOLESKY) is an
 * external, un-vendored, un-versioned dependency that is absent offline and the
 * reference ships no tests or golden numbers for it: for that part PARITY IS UNPINNED
 * (restated from Ceres 2.x semantics).
 */
#ifndef DCS_ORACLE_H
#define DCS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_problem {
  int32_t n_poses, n_edges;
  const double* pose_xyt;  /* N x 3 */
  const int32_t* edge_a;
  const int32_t* edge_b;
  const double* meas_xyt;  /* E x 3 */
  const uint8_t* kind;     /* 0 odometry, 1 closure, 2 bogus */
  int32_t fixed_pose;
  int32_t dcs_on;
  double phi;              /* 0.5 */
  double huber_delta;      /* 0.01 */
  int32_t num_threads;     /* 1 = what the reference ships (Ceres default) */
} oracle_problem;

typedef struct oracle_lm_options {
  int32_t max_num_iterations;
  double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
  double min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  int32_t max_num_consecutive_invalid_steps;
  int32_t jacobi_scaling;
  int32_t verbose;
} oracle_lm_options;

typedef struct oracle_iteration {
  int32_t iteration, step_is_valid, step_is_successful, pad;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm;
  double relative_decrease, trust_region_radius, iteration_time_s, cumulative_time_s;
} oracle_iteration;

typedef struct oracle_summary {
  double initial_cost, final_cost;
  int32_t num_iterations, num_successful_steps, num_unsuccessful_steps, termination_type;
  double total_time_s, eval_time_s, linear_solver_time_s;
  int64_t factor_nnz;
  char message[128];
} oracle_summary;

void oracle_lm_options_default(oracle_lm_options* o);

/* Jet<6> evaluation of the reference functors + Huber corrector, per edge.
 * residuals E x 3, jacobians E x 18 (3x6 row-major: cols = P1[0..2], P2[0..2]),
 * psi E, rho1 E, gradient N x 3.  raw=1 returns residual/Jacobian before the loss
 * corrector (what AutoDiffCostFunction::Evaluate returns). */
int oracle_evaluate(const oracle_problem* p, const double* pose_xyt, int raw, double* cost,
                    double* residuals, double* jacobians, double* psi, double* rho1,
                    double* gradient);

/* Cost-only evaluation with plain doubles (what Ceres does at the candidate point). */
int oracle_cost(const oracle_problem* p, const double* pose_xyt, double* cost);

/* Closed-form (analytic) residual/Jacobian of ONE edge — independent second derivation
 * used to cross-check the Jet path.  e[3], J[18]. dcs: apply the DCS functor. */
void oracle_edge_closed_form(const double* pa, const double* pb, const double* meas, int dcs,
                             double phi, double* e, double* J, double* psi);

/* METHOD 2 (switchable constraints, src/ceres_error.cpp:199-317) — per-edge arithmetic only, pinned against the
 * reference functor; the minimiser for METHOD 2 is not restated yet.
 * e[3] = s * e_plain; J[21] = 3x7 row-major, columns P1[0..2], P2[0..2], s.  Uncorrected (no loss). */
void oracle_sc_edge(const double* pa, const double* pb, const double* meas, double s, double* e, double* J);
/* prior row sqrt(lambda) * (1 - s): e[1], J[1]. */
void oracle_sc_prior(double lambda, double s, double* e, double* J);

/* Upper block pattern {(i,i)} U {(min,max)} over non-constant, touched poses (CSR over
 * poses).  Pass NULL arrays to query nnzb. */
int oracle_pattern(const oracle_problem* p, int32_t* nnzb, int32_t* row_ptr, int32_t* col_idx);

/* H = J^T J (corrected J) on that pattern, nnzb x 9 row-major 3x3 blocks; g = J^T r. */
int oracle_hessian(const oracle_problem* p, const double* pose_xyt, double* block_values,
                   double* gradient);

/* Time `repeats` passes of evaluate + J^T J / J^T r assembly; returns seconds per pass. */
double oracle_time_linearize(const oracle_problem* p, int repeats);

/* Exact solve of (H + diag(lambda)) w = rhs by sparse Cholesky (H from pose_xyt). */
int oracle_linear_solve(const oracle_problem* p, const double* pose_xyt, const double* lambda,
                        const double* rhs, double* w);

/* Ceres-default trust-region LM with an exact sparse normal-Cholesky solve. */
int oracle_solve(const oracle_problem* p, const oracle_lm_options* o, double* pose_xyt_inout,
                 oracle_summary* s, oracle_iteration* trace, int32_t trace_cap);

/* METHOD 2 (switchable constraints, main.cpp:105-150 with SC_ON): Ceres-default LM over poses and one switch per
 * loop edge.  The switches are eliminated edge by edge inside the linear solve (exact Schur complement).
 * switches_inout[E]: entries of odometry edges are ignored; the reference starts every switch at 1.0, lambda = 1.0.
 * p->dcs_on must be 0.  PARITY: functors pinned (oracle_sc_edge), minimiser unpinned like oracle_solve; the
 * elimination is cross-checked against a dense full-system LM in tests/test_oracle_cpu.py. */
int oracle_sc_solve(const oracle_problem* p, double lambda, const oracle_lm_options* o, double* pose_xyt_inout,
                    double* switches_inout, oracle_summary* s, oracle_iteration* trace, int32_t trace_cap);

#ifdef __cplusplus
}
#endif
#endif

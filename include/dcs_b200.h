/*
 * dcs_b200.h — C-ABI of the B200-native DCS-LM hot path (libdcs_b200.so).
 *
 * This is the drop-in boundary for the part of the reference that runs inside
 * ceres::Solve.  The reference has no FFI of its own: its boundary is the Ceres
 * call surface that DCS-ceres/main.cpp touches.  Each entry point below names the
 * reference call(s) it replaces (paths relative to the reference checkout):
 *
 *   dcs_create      <- ceres::Problem ctor (main.cpp:66), new HuberLoss(0.01) (:68),
 *                      OdometryResidue::Create / DCSClosureResidue::Create
 *                      (:98,:113,:127,:136,:147; src/ceres_error.cpp:28-39,121-132),
 *                      Problem::AddResidualBlock (:99,:114,:128,:137,:148),
 *                      SetParameterBlockConstant (:153), Solver::Options (:154-156)
 *   dcs_evaluate    <- Problem::Evaluate / ResidualBlock::Evaluate of the two functors
 *                      (src/ceres_error.cpp:42-94, :135-196) + HuberLoss Corrector
 *   dcs_solve       <- ceres::Solve(options,&problem,&summary) (main.cpp:163)
 *   dcs_get_pattern <- the block structure Ceres' SPARSE_NORMAL_CHOLESKY builds for J^T J
 *                      (main.cpp:156)
 *   dcs_destroy     <- ~Problem
 *
 * Plain pointers and sizes only; every function returns an int status (0 = ok); no
 * exceptions cross the boundary.  All array arguments are HOST pointers unless the
 * name ends in _dev.  There is no CPU fallback: every compute entry point fails with
 * DCS_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef DCS_B200_H
#define DCS_B200_H

#include <stdint.h>

/* The library is built with -fvisibility=hidden: exactly the entry points declared here are exported. */
#if defined(__GNUC__)
#define DCS_API __attribute__((visibility("default")))
#else
#define DCS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum {
  DCS_OK = 0,
  DCS_ERR_ARG = 1,      /* null pointer, bad index, a == b edge, ...            */
  DCS_ERR_CUDA = 2,     /* CUDA runtime error / no device (no CPU fallback)     */
  DCS_ERR_NCCL = 3,     /* NCCL missing or failed (multi-rank handles only)     */
  DCS_ERR_NUMERIC = 4   /* non-finite cost at the initial point                 */
};

/* Edge kinds — reference include/g2o_util.h:13-15 */
enum { DCS_EDGE_ODOMETRY = 0, DCS_EDGE_CLOSURE = 1, DCS_EDGE_BOGUS = 2 };

/* Termination — mirrors ceres::TerminationType as used by the default minimizer */
enum {
  DCS_CONVERGENCE = 0,
  DCS_NO_CONVERGENCE = 1,
  DCS_FAILURE = 2
};

/* Flattened view of the reference's Node / Edge pointer graph (include/graph.h:4-56).
 * Information matrices are not part of it: METHOD 0/1 never read them. */
typedef struct dcs_graph {
  int32_t n_poses;
  int32_t n_edges;
  const double* pose_xyt;   /* n_poses x 3, row-major (x, y, theta); copied in          */
  const int32_t* edge_a;    /* first endpoint  (Edge::a->index)                          */
  const int32_t* edge_b;    /* second endpoint (Edge::b->index)                          */
  const double* meas_xyt;   /* n_edges x 3 (Edge::x, y, theta)                           */
  const uint8_t* kind;      /* DCS_EDGE_*; order = odometry, closure, bogus (main.cpp)   */
  int32_t fixed_pose;       /* parameter block held constant (main.cpp:153 -> 0); -1: none */
} dcs_graph;

typedef struct dcs_options {
  /* residual model */
  int32_t dcs_on;                 /* METHOD==1 (main.cpp:55)                             */
  double phi;                     /* 0.5  (src/ceres_error.cpp:185)                      */
  double huber_delta;             /* 0.01 (main.cpp:68)                                  */
  /* trust-region LM — Ceres defaults implied by main.cpp:154-156                         */
  int32_t max_num_iterations;     /* 50                                                  */
  double initial_trust_region_radius; /* 1e4                                             */
  double max_trust_region_radius;     /* 1e16                                            */
  double min_trust_region_radius;     /* 1e-32                                           */
  double min_relative_decrease;       /* 1e-3                                            */
  double min_lm_diagonal;             /* 1e-6                                            */
  double max_lm_diagonal;             /* 1e32                                            */
  double function_tolerance;          /* 1e-6                                            */
  double gradient_tolerance;          /* 1e-10                                           */
  double parameter_tolerance;         /* 1e-8                                            */
  int32_t max_num_consecutive_invalid_steps; /* 5                                        */
  int32_t jacobi_scaling;             /* 1                                               */
  /* linear solver: preconditioned CG on the 3x3-block normal equations (see `preconditioner`) */
  double pcg_rel_tol;             /* stop when |r|_2 <= pcg_rel_tol * |rhs|_2            */
  int32_t pcg_max_iter;
  int32_t pcg_check_every;        /* iterations between convergence tests (one CUDA-graph launch; on graphs of <= 8192
                                   * poses the whole solve is one kernel and the test runs inside it at this period) */
  int32_t preconditioner;         /* 0: 3x3 block-Jacobi; 1 (default): block-Jacobi over chain segments of 32 poses
                                   * (each block = the block-tridiagonal odometry-chain part of the segment,
                                   * factorised exactly once per LM iteration)                                    */
  /* execution */
  int32_t device;                 /* CUDA device ordinal                                 */
  int32_t verbose;                /* 1: Ceres-style progress table on stdout             */
  /* multi-GPU: one process per GPU (world <= 8: the GPUs of one NVSwitch node). world==1 -> single device.
   * With world > 1 every call that takes or returns poses (dcs_create, dcs_evaluate, dcs_linearize, dcs_cost,
   * dcs_solve) is COLLECTIVE: all ranks make the same calls in the same order with the same full-size arrays;
   * a rank reads only its own pose rows (dcs_partition) from the input and receives its halo from the owners
   * over NVLink (peer-memory push; DCS_HALO=nccl forces grouped ncclSend/ncclRecv).  Outputs: scalars (cost,
   * summaries, traces) and the poses written by dcs_solve are complete and identical on every rank; per-row
   * vector outputs (the gradient of dcs_evaluate / dcs_linearize, w of dcs_pcg_solve) are filled for the
   * calling rank's own rows only and zero elsewhere (sum over ranks = the full vector).              */
  int32_t rank;
  int32_t world;
  const void* nccl_unique_id;     /* 128-byte ncclUniqueId shared by all ranks           */
  double max_solver_time_s;       /* 1e6: Solver::Options::max_solver_time_in_seconds (checked between iterations; with
                                   * world > 1 rank 0's clock decides for everybody)      */
  /* METHOD 2 — switchable constraints (main.cpp:55 SC_ON, :105-150; src/ceres_error.cpp:203-317): every closure / bogus
   * edge carries a scalar switch s (initial 1), e = s * e_plain with HuberLoss(0.01), plus the prior row
   * sqrt(lambda) (1 - s) without loss.  Exclusive with dcs_on; single-rank handles only.  The switches are
   * eliminated edge by edge inside the linear solve (exact Schur complement onto the 3x3 pose blocks) and
   * recovered after it; dcs_get_switches returns them.                                                        */
  int32_t switchable_on;          /* METHOD==2                                           */
  double switch_prior_lambda;     /* 1.0 (main.cpp:110 sc_prior_lambda)                  */
} dcs_options;

typedef struct dcs_iteration {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_successful;
  int32_t linear_solver_iterations;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  double linear_solver_residual;  /* final PCG |r|/|rhs|                                 */
  double iteration_time_s;
  double cumulative_time_s;
  double linear_solver_true_residual; /* |(H + Lambda) w - g| / |g| recomputed in fp64 from the returned step w
                                       * (independent of the PCG recurrence); what replaces "the factorisation is
                                       * exact" at sizes no direct solver reaches                                */
} dcs_iteration;

typedef struct dcs_summary {
  double initial_cost;
  double final_cost;
  int32_t num_iterations;         /* entries written to the trace (iteration 0 included; as in Ceres, the iteration
                                   * that ends on the parameter / function tolerance is not appended)          */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t termination_type;       /* DCS_CONVERGENCE / DCS_NO_CONVERGENCE / DCS_FAILURE  */
  int64_t total_pcg_iterations;
  double total_time_s;
  double eval_time_s;             /* device time in eval+assembly launches               */
  double linear_solver_time_s;    /* device time in PCG                                  */
  char message[128];
} dcs_summary;

typedef struct dcs_handle dcs_handle;

/* Fill o with the reference's defaults (see field comments). */
DCS_API void dcs_options_default(dcs_options* o);

/* Library / device probes (no compute). */
DCS_API const char* dcs_version(void);
DCS_API int dcs_device_count(void);

/* Pose-range / edge-slice partition a handle with (rank, world) uses: contiguous, equal-sized pose ranges
 * following the odometry chain (padded to the kernel's row window), equal edge slices for the cost-only
 * kernel.  Pure host arithmetic (no CUDA): out = {row_lo, n_rows, rows_per_rank, edge_lo, edge_hi}. */
DCS_API int dcs_partition(int32_t n_poses, int32_t n_edges, int32_t rank, int32_t world, int32_t out[5]);

/* 128-byte id for a multi-rank group; rank 0 calls it and ships the bytes to the peers. */
DCS_API int dcs_nccl_unique_id(void* out128);

/* Upload the graph, build the half-edge CSR and the block pattern (one-time sort). */
DCS_API int dcs_create(const dcs_graph* g, const dcs_options* o, dcs_handle** out);
DCS_API void dcs_destroy(dcs_handle* h);

/* Parity hook: evaluate at pose_xyt (NULL -> the handle's current poses).
 * Any output pointer may be NULL.  Shapes: residuals E x 3, jacobians E x 18
 * (row-major 3x6 per edge: d e / d(pa, pb), after DCS and the Huber corrector),
 * psi E, rho1 E (Huber rho'), gradient N x 3 (zeros at the fixed pose). */
/* Single-rank METHOD 0/1 handles for the per-edge outputs (cost and gradient work everywhere; with switchable_on the
 * cost includes the switch priors and the gradient is the pose part at the current switches). */
DCS_API int dcs_evaluate(dcs_handle* h, const double* pose_xyt, double* cost,
                 double* residuals, double* jacobians, double* psi, double* rho1,
                 double* gradient);

/* Hot path on host buffers: H2D poses, fused eval + J^T J / J^T r assembly on device,
 * D2H cost and gradient (either may be NULL).  The assembled H stays on the device. */
DCS_API int dcs_linearize(dcs_handle* h, const double* pose_xyt, double* cost, double* gradient);

/* Same launches with everything resident on the device (used by bench.py's `value`): `repeats` times the fused
 * eval + assembly launch (+ its scalar fold), device time in ms.  The assembled product is the reference's structure
 * (one block per edge + diagonal blocks + gradient).  with_solver_setup != 0 also times the expansion of the edge
 * blocks into the row storage the PCG's SpMV walks (both triangles; the linear solver's setup, paid once per LM
 * iteration). */
DCS_API int dcs_linearize_resident(dcs_handle* h, int32_t repeats, int32_t with_solver_setup, float* ms_total);

/* Cost-only evaluation (candidate point inside LM). */
DCS_API int dcs_cost(dcs_handle* h, const double* pose_xyt, double* cost);

/* Integer parity hook: upper block pattern of J^T J over the non-constant poses,
 * {(i,i)} U {(min(a,b),max(a,b))}, as CSR over poses.  Call with NULL arrays to get
 * the sizes; then with row_ptr[n_poses+1] and col_idx[nnzb]. */
DCS_API int dcs_get_pattern(dcs_handle* h, int32_t* n_block_rows, int32_t* nnzb,
                    int32_t* row_ptr, int32_t* col_idx);

/* Values of the assembled J^T J on that pattern (nnzb x 9, row-major 3x3), from the last
 * dcs_linearize / dcs_evaluate / accepted LM step. */
DCS_API int dcs_get_hessian(dcs_handle* h, double* block_values);

/* One linear solve (H + diag(lambda)) w = rhs with block-Jacobi PCG on the current H.
 * lambda, rhs, w: N x 3 host arrays (entries of the fixed pose ignored / zero). */
DCS_API int dcs_pcg_solve(dcs_handle* h, const double* lambda, const double* rhs, double* w,
                  int32_t* iterations, double* rel_residual);

/* Full DCS-LM solve. pose_xyt_inout: N x 3, updated in place (Node::p write-back,
 * include/graph.h:10-17).  trace may be NULL; trace_cap entries are written at most. */
DCS_API int dcs_solve(dcs_handle* h, double* pose_xyt_inout, dcs_summary* summary,
              dcs_iteration* trace, int32_t trace_cap);

/* Batched tiny solves — the form the METHOD 3/4 clients need (src/layer_manager.cpp:137-179, :602-654;
 * src/simple_layer_manager.cpp:173-211, :567-622: a fresh ceres::Problem of odometry + a few candidate loop edges with
 * OdometryResidue + HuberLoss, 1-2 iterations of ceres::Solve, summary.final_cost read back, thousands of times).
 * Every item is an independent problem (its own graph, poses copied in; pose_xyt_inout may be NULL when only the
 * summary is wanted); all items run with the same options.  n_threads host threads (<= 0: 8) each drive one
 * handle / stream at a time, so the launches of several small solves overlap on the device.  status per item;
 * the call returns the first non-zero status.  Results are identical to n_items separate dcs_create / dcs_solve
 * / dcs_destroy sequences (every solve is deterministic on its own stream). */
typedef struct dcs_batch_item {
  dcs_graph graph;
  double* pose_xyt_inout;         /* n_poses x 3 or NULL                                 */
  dcs_summary summary;            /* out                                                 */
  int32_t status;                 /* out: DCS_OK or the error of this item               */
} dcs_batch_item;
DCS_API int dcs_solve_batch(dcs_batch_item* items, int32_t n_items, const dcs_options* options, int32_t n_threads);

/* METHOD 2: the switch value of every edge after the last dcs_solve (n_edges doubles, residual-block order;
 * odometry edges carry no switch and read 1.0) — what main.cpp:169-171 hands to writePoseGraph_switches. */
DCS_API int dcs_get_switches(dcs_handle* h, double* switches);

/* Page-locked host buffers (cudaHostAlloc).  Pose / gradient arrays handed to dcs_linearize, dcs_evaluate and
 * dcs_solve may live anywhere; when they are page-locked the library DMAs straight from / into them instead of
 * staging through its own pinned bounce buffer. */
DCS_API void* dcs_host_alloc(uint64_t bytes);
DCS_API void dcs_host_free(void* p);

/* Last CUDA / NCCL error text for this thread ("" if none). */
DCS_API const char* dcs_last_error(void);

/* Number of kernel launches issued by this library since the counter was last reset. */
DCS_API int64_t dcs_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* DCS_B200_H */

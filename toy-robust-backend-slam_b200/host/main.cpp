// main.cpp — drop-in for the reference's DCS-ceres/main.cpp, METHOD 0 (baseline), 1 (DCS) and 2 (switchable constraints).
//
//   ./main DATASET_NAME_WITHOUT_DOTG2O NUM_OUTLIER_LOOPS METHOD
//
// Same positionals, same stdout lines in the same order, same ../data and ../save paths and
// the same four output files for drawer/ as the reference (main.cpp:32-173).  The block the
// reference spends in Ceres (main.cpp:66-164: Problem, HuberLoss, AddResidualBlock per edge,
// SetParameterBlockConstant, Solve, FullReport) is ONE call sequence into the CUDA library:
// dcs_create -> dcs_solve -> dcs_destroy (include/dcs_b200.h).  There is no CPU fallback:
// without a usable B200 the program reports the CUDA error and exits non-zero.
//
// METHOD 2 adds one switch + prior per loop edge (main.cpp:105-150) and writes save/switches.txt (main.cpp:169-171).
// METHOD 3/4 (layer managers) are outside this path; they are refused with a message instead of silently
// running something else.
//
// Extra knobs, all through the environment so the CLI stays identical:
//   DCS_SEED         seed for the outlier injection (reference: time(0), main.cpp:43)
//   DCS_DATA_PATH    overrides ../data        DCS_SAVE_PATH   overrides ../save
//   DCS_PCG_TOL      PCG relative tolerance   DCS_PCG_MAX_ITER
//   DCS_DEVICE       CUDA device ordinal
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <string>
#include <vector>

#include "dcs_b200.h"
#include "g2o_util.h"
#include "graph.h"

using std::cout;
using std::endl;
using std::string;

static string env_or(const char* name, const string& dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? string(v) : dflt;
}

static const char* termination_name(int t) {
  return t == DCS_CONVERGENCE ? "CONVERGENCE" : t == DCS_NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE";
}

int main(int argc, char* argv[]) {
  if (argc < 4) {
    cout << "Usage: " << argv[0] << " DATASET NUM_OUTLIER_LOOPS METHOD\n";
    cout << "METHOD: 0=baseline, 1=DCS, 2=Switchable (3=Layer, 4=Simple Layer MCTS: not on this path)\n";
    cout << "Example: " << argv[0] << " INTEL 50 1\n";
    return -1;
  }
  const string BASE_PATH = env_or("DCS_DATA_PATH", "../data");
  const string SAVE_PATH = env_or("DCS_SAVE_PATH", "../save");

  const char* seed_env = std::getenv("DCS_SEED");
  std::srand(seed_env && *seed_env ? (unsigned int)std::strtoul(seed_env, nullptr, 10) : (unsigned int)time(0));

  const string fpath = BASE_PATH + "/" + string(argv[1]) + ".g2o";
  cout << "Start Reading PoseGraph\n";
  ReadG2O g2o_manager;
  if (!g2o_manager.read(fpath)) std::cerr << "cannot open " << fpath << endl;

  const int num_bogus_loops = atoi(argv[2]);
  g2o_manager.add_random_C(num_bogus_loops);

  const int METHOD = atoi(argv[3]);
  if (METHOD != 0 && METHOD != 1 && METHOD != 2) {
    std::cerr << "METHOD " << METHOD << " is outside the B200 hot path (only 0=baseline, 1=DCS, 2=Switchable)." << endl;
    return 2;
  }

  g2o_manager.writePoseGraph_nodes(SAVE_PATH + "/init_nodes.txt");
  g2o_manager.writePoseGraph_edges(SAVE_PATH + "/init_edges.txt");
  cout << "total nodes : " << g2o_manager.nNodes.size() << endl;
  cout << "total nEdgesOdometry : " << g2o_manager.nEdgesOdometry.size() << endl;
  cout << "total nEdgesClosure : " << g2o_manager.nEdgesClosure.size() << endl;
  cout << "total nEdgesBogus : " << g2o_manager.nEdgesBogus.size() << endl;
  if (g2o_manager.nNodes.empty()) return 1;

  // Problem + HuberLoss(0.01) + one residual block per edge + constant first pose + default
  // trust-region options, as one flat graph.
  FlatGraph flat;
  g2o_manager.flatten(&flat);
  dcs_graph graph;
  graph.n_poses = (int32_t)g2o_manager.nNodes.size();
  graph.n_edges = (int32_t)flat.kind.size();
  graph.pose_xyt = flat.pose_xyt.data();
  graph.edge_a = flat.edge_a.data();
  graph.edge_b = flat.edge_b.data();
  graph.meas_xyt = flat.meas_xyt.data();
  graph.kind = flat.kind.data();
  graph.fixed_pose = 0;  // problem.SetParameterBlockConstant(nNodes[0]->p)

  dcs_options options;
  dcs_options_default(&options);
  options.dcs_on = (METHOD == 1);
  options.switchable_on = (METHOD == 2);     // SC_ON (main.cpp:55)
  options.switch_prior_lambda = 1.0;         // sc_prior_lambda (main.cpp:110)
  options.verbose = 1;  // minimizer_progress_to_stdout
  options.device = atoi(env_or("DCS_DEVICE", "0").c_str());
  if (std::getenv("DCS_PCG_TOL")) options.pcg_rel_tol = atof(std::getenv("DCS_PCG_TOL"));
  if (std::getenv("DCS_PCG_MAX_ITER")) options.pcg_max_iter = atoi(std::getenv("DCS_PCG_MAX_ITER"));

  dcs_handle* handle = nullptr;
  int rc = dcs_create(&graph, &options, &handle);
  if (rc != DCS_OK) {
    std::cerr << "dcs_create failed (" << rc << "): " << dcs_last_error() << endl;
    return 3;
  }
  dcs_summary summary;
  std::vector<dcs_iteration> trace((size_t)options.max_num_iterations + 2);
  rc = dcs_solve(handle, flat.pose_xyt.data(), &summary, trace.data(), (int32_t)trace.size());
  if (rc != DCS_OK) {
    std::cerr << "dcs_solve failed (" << rc << "): " << dcs_last_error() << endl;
    dcs_destroy(handle);
    return 3;
  }
  g2o_manager.scatter_poses(flat.pose_xyt.data());  // Ceres mutates Node::p in place

  // summary.FullReport() stand-in
  cout << "\nSolver Summary (B200 DCS-LM, block-Jacobi PCG)\n\n";
  cout << "Parameter blocks   " << graph.n_poses << " (1 constant)\n";
  cout << "Residual blocks    " << graph.n_edges << "\n";
  cout << "Robust loss        HuberLoss(0.01)" << (options.dcs_on ? " + DCS(phi=0.5) on loop edges" : "")
       << (options.switchable_on ? " + one switch and prior (lambda=1) per loop edge" : "") << "\n\n";
  std::printf("Cost:\nInitial        %.6e\nFinal          %.6e\nChange         %.6e\n\n", summary.initial_cost,
              summary.final_cost, summary.initial_cost - summary.final_cost);
  std::printf("Minimizer iterations   %d\nSuccessful steps       %d\nUnsuccessful steps     %d\n", summary.num_iterations,
              summary.num_successful_steps, summary.num_unsuccessful_steps);
  std::printf("PCG iterations (total) %lld\n\n", (long long)summary.total_pcg_iterations);
  std::printf("Time (in seconds):\n  Residual+Jacobian+assembly (device)  %.6f\n  Linear solver (device)               %.6f\n  Total                                %.6f\n\n",
              summary.eval_time_s, summary.linear_solver_time_s, summary.total_time_s);
  cout << "Termination:   " << termination_name(summary.termination_type) << " (" << summary.message << ")\n" << endl;

  g2o_manager.writePoseGraph_nodes(SAVE_PATH + "/opt_nodes.txt");
  g2o_manager.writePoseGraph_edges(SAVE_PATH + "/opt_edges.txt");
  if (METHOD == 2) {   // main.cpp:169-171: priors (all 1.0) and optimised switches, closure edges first, then bogus
    std::vector<double> all((size_t)graph.n_edges, 1.0);
    rc = dcs_get_switches(handle, all.data());
    if (rc != DCS_OK) {
      std::cerr << "dcs_get_switches failed (" << rc << "): " << dcs_last_error() << endl;
      dcs_destroy(handle);
      return 3;
    }
    const size_t n_odo = g2o_manager.nEdgesOdometry.size();
    const std::vector<double> optimized(all.begin() + (long)n_odo, all.end());
    const std::vector<double> priors(optimized.size(), 1.0);
    g2o_manager.writePoseGraph_switches(SAVE_PATH + "/switches.txt", priors, optimized);
  }
  dcs_destroy(handle);
  return 0;
}

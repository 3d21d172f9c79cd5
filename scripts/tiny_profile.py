"""Where the time of a tiny solve goes (N3): handle creation vs the 2-iteration solve, INTEL-sized variants."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
gi = D.Graph.load_npz(os.path.join(ROOT, "tests", "golden", "INTEL_50_seed1.npz"))
rng = np.random.default_rng(5)
odo, loops = np.flatnonzero(gi.kind == 0), np.flatnonzero(gi.kind != 0)
keep = np.r_[odo, np.sort(rng.choice(loops, size=180, replace=False))]
g = D.Graph(gi.pose_xyt, gi.edge_a[keep], gi.edge_b[keep], gi.meas_xyt[keep], gi.kind[keep])
D.Solver(g, dcs_on=False, max_num_iterations=2).close()
tc = ts = 0.0; n = 20
for i in range(n):
    t = time.perf_counter(); s = D.Solver(g, dcs_on=False, max_num_iterations=2); tc += time.perf_counter() - t
    t = time.perf_counter(); x, sm, tr = s.solve(); ts += time.perf_counter() - t
    s.close()
print(f"create {1e3 * tc / n:.2f} ms, solve(2 iterations) {1e3 * ts / n:.2f} ms, pcg iterations {sm.total_pcg_iterations}, "
      f"linear solver {1e3 * sm.linear_solver_time_s:.2f} ms, eval {1e3 * sm.eval_time_s:.3f} ms, us/pcg {1e6 * sm.linear_solver_time_s / max(1, sm.total_pcg_iterations):.1f}")
for it in tr: print(it.iteration, it.cost, it.linear_solver_iterations)

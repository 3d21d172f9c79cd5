"""Short driver for ncu captures: 1M-pose synthetic graph, a few fused eval+assembly launches and a short PCG."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
s = D.Solver(g, dcs_on=True, max_num_iterations=1, pcg_max_iter=32, pcg_check_every=8, pcg_rel_tol=1e-30)
print("linearize ms", s.linearize_resident(5) / 5)
x, sm, tr = s.solve()
print("pcg iters", sm.total_pcg_iterations, "us/iter", 1e6 * sm.linear_solver_time_s / max(1, sm.total_pcg_iterations))

"""Full DCS-LM solve (BASELINE metric 2: LM iterations/s) on the synthetic graph: per-iteration trace as JSON lines.
usage: lm_full.py [n_poses] [max_iter] [pcg_rel_tol] [max_seconds]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-12
extra = {}
if len(sys.argv) > 4 and hasattr(D.Options, "max_solver_time_s"):
    extra["max_solver_time_s"] = float(sys.argv[4])
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
t = time.time()
s = D.Solver(g, dcs_on=True, max_num_iterations=iters, pcg_rel_tol=tol, **extra)
print(json.dumps({"create_s": time.time() - t, "n_poses": N, "n_edges": g.n_edges}), flush=True)
t = time.time()
x, sm, tr = s.solve()
dt = time.time() - t
for it in tr:
    print(json.dumps({k: getattr(it, k) for k, _ in it._fields_}), flush=True)
print(json.dumps({"seconds": dt, "iterations": sm.num_iterations - 1, "lm_it_per_s": (sm.num_iterations - 1) / dt,
                  "pcg_iterations": int(sm.total_pcg_iterations), "eval_s": sm.eval_time_s, "pcg_s": sm.linear_solver_time_s,
                  "initial_cost": sm.initial_cost, "final_cost": sm.final_cost, "termination": sm.termination_type,
                  "message": sm.message.decode()}), flush=True)

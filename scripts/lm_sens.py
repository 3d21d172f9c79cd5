"""Sensitivity of the 50-iteration DCS-LM trajectory at 1 M poses to the accuracy / rounding of the linear solves:
the same solve with pcg_rel_tol 1e-12 (default), 1e-13 and with the other preconditioner (different rounding path).
Prints, per variant, the iteration-by-iteration cost and accept flag; the comparison is done offline."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
out = {}
for tag, opts in (("tol1e-12", dict(pcg_rel_tol=1e-12)), ("tol1e-13", dict(pcg_rel_tol=1e-13)), ("tol1e-12_blockjacobi", dict(pcg_rel_tol=1e-12, preconditioner=0))):
    t = time.time()
    with D.Solver(g, dcs_on=True, **opts) as s:
        x, sm, tr = s.solve()
    out[tag] = dict(seconds=time.time() - t, final_cost=sm.final_cost, pcg_iterations=int(sm.total_pcg_iterations),
                    cost=[t.cost for t in tr], ok=[t.step_is_successful for t in tr], radius=[t.trust_region_radius for t in tr],
                    true_res=[t.linear_solver_true_residual for t in tr], pcg=[t.linear_solver_iterations for t in tr])
    print(tag, out[tag]["seconds"], sm.final_cost, sm.total_pcg_iterations, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lm_sens_r02.json"), "w"))

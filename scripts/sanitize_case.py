"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): create, evaluate, pattern, 3 LM iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import numpy as np, dcs_b200 as D
g = D.Graph.load_npz(os.path.join(ROOT, "tests", "golden", "INTEL_50_seed1.npz"))
for pc in (0, 1):
    with D.Solver(g, dcs_on=True, max_num_iterations=3, preconditioner=pc, pcg_max_iter=256) as s:
        ev = s.evaluate()
        rp, ci, hv = s.hessian()
        x, sm, tr = s.solve()
        print("precond", pc, "cost", ev["cost"], "nnzb", ci.size, "final", sm.final_cost, "pcg", sm.total_pcg_iterations)
g2 = D.Graph.synthetic(3000, 8100, n_bogus=900)
with D.Solver(g2, dcs_on=True, max_num_iterations=2, pcg_max_iter=128) as s:
    print("synthetic", s.solve()[1].final_cost)

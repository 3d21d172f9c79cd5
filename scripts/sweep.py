"""BASELINE.json configs[4]: synthetic scaling sweep (1 GPU leg): eval+assembly and PCG iteration time vs graph size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
sizes = [int(float(x)) for x in (sys.argv[1:] or ["1e5", "3e5", "1e6", "3e6", "1e7"])]
out = []
for N in sizes:
    t = time.time(); g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N)); tg = time.time() - t
    t = time.time(); s = D.Solver(g, dcs_on=True, max_num_iterations=1, pcg_max_iter=320, pcg_check_every=32, pcg_rel_tol=1e-30); tc = time.time() - t
    s.linearize_resident(5)
    us = 1e3 * s.linearize_resident(20) / 20
    x, sm, tr = s.solve()
    rec = dict(n_poses=N, n_edges=g.n_edges, gen_s=round(tg, 2), create_s=round(tc, 3), linearize_us=round(us, 1),
               edges_per_s=g.n_edges / us * 1e6, frac=(108 * g.n_edges + 120 * N) / us / 1e3 / 6444.4,
               pcg_us_per_iter=round(1e6 * sm.linear_solver_time_s / max(1, sm.total_pcg_iterations), 1))
    print(json.dumps(rec), flush=True)
    out.append(rec)
    s.close(); del g
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep_r01.json"), "w"), indent=1)

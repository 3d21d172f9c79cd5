"""dcs_solve_batch throughput (64 INTEL-sized variants x 2 LM iterations, handle creation included) with the cluster PCG
(default) and with the general path (DCS_PCG_CLUSTER=0), 1 / 8 / 16 host threads; and create / solve time of one item."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
gi = D.Graph.load_npz(os.path.join(ROOT, "tests", "golden", "INTEL_50_seed1.npz"))
rng = np.random.default_rng(5)
odo, loops = np.flatnonzero(gi.kind == 0), np.flatnonzero(gi.kind != 0)
variants = []
for v in range(64):
    keep = np.r_[odo, np.sort(rng.choice(loops, size=150 + v, replace=False))]
    variants.append(D.Graph(gi.pose_xyt, gi.edge_a[keep], gi.edge_b[keep], gi.meas_xyt[keep], gi.kind[keep]))
out = {}
for label, env in (("cluster_pcg", "1"), ("general_pcg", "0")):
    os.environ["DCS_PCG_CLUSTER"] = env
    D.Solver(variants[0], dcs_on=False, max_num_iterations=2).close()
    tc = ts = 0.0; n = 20
    for i in range(n):
        t = time.perf_counter(); s = D.Solver(variants[i], dcs_on=False, max_num_iterations=2); tc += time.perf_counter() - t
        t = time.perf_counter(); x, sm, tr = s.solve(); ts += time.perf_counter() - t
        s.close()
    r = {"create_ms": 1e3 * tc / n, "solve_ms": 1e3 * ts / n, "pcg_iterations": int(sm.total_pcg_iterations)}
    for nt in (1, 8, 16):
        D.solve_batch(variants[:4], dcs_on=False, n_threads=nt, max_num_iterations=2)
        t0 = time.perf_counter()
        sums, _ = D.solve_batch(variants, dcs_on=False, n_threads=nt, max_num_iterations=2)
        r[f"solves_per_sec_threads_{nt}"] = len(variants) / (time.perf_counter() - t0)
    r["final_cost_first"] = sums[0].final_cost
    out[label] = r
print(json.dumps(out))

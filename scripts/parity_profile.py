"""Per-iteration deviation of the GPU LM traces from the committed oracle traces (cost, radius, accept flag):
SYN10K (+1000 outliers, two injection seeds, METHOD 1) and METHOD 2 on INTEL+50 / M3500+100.  -> gpurun_out/parity_profile_r02.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
GOLD = os.path.join(ROOT, "tests", "golden")
out = {}


def profile(tag, tr, sm, co, ro, oko, fco, extra=None):
    n = min(len(tr), len(co))
    cg = np.array([t.cost for t in tr[:n]]); rg = np.array([t.trust_region_radius for t in tr[:n]])
    okg = [int(t.step_is_successful) for t in tr[:n]]
    same = [a == int(b) for a, b in zip(okg, oko[:n])]
    rec = dict(n_gpu=len(tr), n_oracle=len(co), first_accept_mismatch=(same.index(False) if False in same else None),
               cost_rel=[float(v) for v in np.abs(cg - co[:n]) / np.abs(co[:n])], radius_rel=[float(v) for v in np.abs(rg - ro[:n]) / np.abs(ro[:n])],
               ok_oracle=[int(v) for v in oko[:n]], final_cost_gpu=sm.final_cost, final_cost_oracle=float(fco),
               final_cost_rel=abs(sm.final_cost - float(fco)) / float(fco), max_true_residual=max(t.linear_solver_true_residual for t in tr),
               pcg_iterations=int(sm.total_pcg_iterations))
    if extra: rec.update(extra)
    out[tag] = rec
    print(tag, "n", rec["n_gpu"], rec["n_oracle"], "mismatch at", rec["first_accept_mismatch"], "final rel %.2e" % rec["final_cost_rel"],
          "max cost rel %.2e" % max(rec["cost_rel"]), "max radius rel %.2e" % max(rec["radius_rel"]), flush=True)
    print("   cost_rel", " ".join("%.0e" % v for v in rec["cost_rel"]), flush=True)


for fx in ("SYN10K_1000_s777", "SYN10K_1000"):
    z = np.load(os.path.join(GOLD, fx + ".npz"))
    g = D.Graph(z["pose_xyt"], z["edge_a"], z["edge_b"], z["meas_xyt"], z["kind"], int(z["fixed_pose"]))
    for tol in (1e-12, 1e-13):
        with D.Solver(g, dcs_on=True, pcg_rel_tol=tol) as s:
            x, sm, tr = s.solve()
        profile(f"{fx}_tol{tol:g}", tr, sm, z["trace_cost_dcs1"], z["trace_radius_dcs1"], z["trace_ok_dcs1"], z["final_cost_dcs1"],
                dict(pose_max_abs=float(np.abs(x - z["final_pose_dcs1"]).max())))
z = np.load(os.path.join(GOLD, "method2_traces.npz"))
for name in ("INTEL_50_seed1", "M3500_100_seed1"):
    g = D.Graph.load_npz(os.path.join(GOLD, name + ".npz"))
    for tol in (1e-12, 1e-13):
        with D.Solver(g, dcs_on=False, switchable_on=1, pcg_rel_tol=tol) as s:
            x, sm, tr = s.solve()
            sw = s.switches()
        loops = g.kind != 0
        profile(f"method2_{name}_tol{tol:g}", tr, sm, z[f"{name}_trace_cost"], z[f"{name}_trace_radius"], z[f"{name}_trace_ok"], z[f"{name}_final_cost"],
                dict(switch_max_abs=float(np.abs(sw[loops] - z[f"{name}_switches"][loops]).max()), pose_max_abs=float(np.abs(x - z[f"{name}_final_pose"]).max())))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_profile_r02.json"), "w"), indent=1)

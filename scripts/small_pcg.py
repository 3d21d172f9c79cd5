"""Small graphs: the one-launch cluster PCG (k_pcg_cluster, default) against the general path (DCS_PCG_CLUSTER=0) on the
reference's datasets: same linear-solve result, same LM trace, time per PCG iteration and per full solve.
Prints one JSON line per case; exits non-zero when the two paths disagree."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D


def solver(g, cluster, **kw):
    os.environ["DCS_PCG_CLUSTER"] = "1" if cluster else "0"     # read by dcs_create
    try:
        return D.Solver(g, **kw)
    finally:
        os.environ.pop("DCS_PCG_CLUSTER", None)


def case(name, dcs_on=True, **kw):
    g = D.Graph.load_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    out = {"case": name, "n_poses": int(g.n_poses), "n_edges": int(g.n_edges), "options": {"dcs_on": dcs_on, **kw}}
    ok = True
    # one linear solve at the initial point, lambda = 1e-4 * diag-scale stand-in
    rng = np.random.default_rng(3)
    rhs = rng.normal(0, 1, (g.n_poses, 3)); lam = np.full((g.n_poses, 3), 1e-3)
    res = {}
    for cl in ((1, 0) if not kw.get("switchable_on") else ()):
        with solver(g, cl, dcs_on=dcs_on, **kw) as s:
            s.linearize(g.pose_xyt)
            s.pcg_solve(lam, rhs)                                    # warm-up (graph capture / module load)
            t = time.perf_counter(); w, it, rel = s.pcg_solve(lam, rhs); dt = time.perf_counter() - t
            res[cl] = (w, it, rel, dt)
    if res:
        w1, it1, rel1, dt1 = res[1]; w0, it0, rel0, dt0 = res[0]
        dw = float(np.abs(w1 - w0).max() / np.abs(w0).max())
        out["linear_solve"] = {"iters_cluster": it1, "iters_general": it0, "rel_residual_cluster": rel1, "rel_residual_general": rel0,
                               "w_rel_diff": dw, "us_per_iter_cluster": 1e6 * dt1 / max(1, it1), "us_per_iter_general": 1e6 * dt0 / max(1, it0)}
        ok &= dw <= 1e-6 and rel1 <= 1.0000001e-12 and abs(it1 - it0) <= 64
    # full LM solves
    tr = {}
    for cl in (1, 0):
        with solver(g, cl, dcs_on=dcs_on, **kw) as s:
            s.solve()                                                # warm-up on a handle of its own
        with solver(g, cl, dcs_on=dcs_on, **kw) as s:
            t = time.perf_counter(); x, sm, trace = s.solve(); dt = time.perf_counter() - t
            tr[cl] = (x, sm, trace, dt)
    x1, sm1, t1, dt1 = tr[1]; x0, sm0, t0, dt0 = tr[0]
    same_seq = len(t1) == len(t0) and all(a.step_is_successful == b.step_is_successful for a, b in zip(t1, t0))
    cost_rel = max(abs(a.cost - b.cost) / abs(b.cost) for a, b in zip(t1, t0)) if same_seq else None
    out["lm"] = {"iterations": len(t1), "same_accept_sequence": bool(same_seq), "max_cost_rel_diff": cost_rel,
                 "final_cost_cluster": sm1.final_cost, "final_cost_general": sm0.final_cost,
                 "pcg_iterations_cluster": int(sm1.total_pcg_iterations), "pcg_iterations_general": int(sm0.total_pcg_iterations),
                 "seconds_cluster": dt1, "seconds_general": dt0,
                 "us_per_pcg_cluster": 1e6 * sm1.linear_solver_time_s / max(1, sm1.total_pcg_iterations),
                 "us_per_pcg_general": 1e6 * sm0.linear_solver_time_s / max(1, sm0.total_pcg_iterations),
                 "max_true_residual_cluster": max(i.linear_solver_true_residual for i in t1)}
    ok &= same_seq and cost_rel is not None and cost_rel <= 1e-8 and abs(sm1.final_cost - sm0.final_cost) <= 1e-9 * abs(sm0.final_cost)
    out["ok"] = bool(ok)
    print(json.dumps(out), flush=True)
    return ok


if __name__ == "__main__":
    good = True
    good &= case("INTEL_50_seed1")
    good &= case("M3500_100_seed1")
    good &= case("INTEL_0_seed1", dcs_on=False)
    good &= case("M3500_100_seed1", dcs_on=False, switchable_on=1)
    sys.exit(0 if good else 1)

"""Development check on a GPU box: parity printouts + first timings (not a test, not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import dcs_b200 as D
import oracle_py as O

print(D.version(), "devices", D.device_count())
for name in ("INTEL_50_seed1", "M3500_100_seed1"):
    z = np.load(f"{ROOT}/tests/golden/{name}.npz")
    g = D.Graph.load_npz(f"{ROOT}/tests/golden/{name}.npz")
    for dcs in (1, 0):
        ora = O.Oracle(g, dcs_on=bool(dcs))
        with D.Solver(g, dcs_on=bool(dcs)) as s:
            ev = s.evaluate(); ref = ora.evaluate()
            print(name, "dcs", dcs, "cost gpu", ev["cost"], "oracle", ref["cost"], "rel", abs(ev["cost"]-ref["cost"])/ref["cost"])
            print("   |dr|", np.abs(ev["residuals"]-ref["residuals"]).max(), "|dJ|", np.abs(ev["jacobians"]-ref["jacobians"]).max(),
                  "|dg|", np.abs(ev["gradient"]-ref["gradient"]).max(), "psi", np.abs(ev["psi"]-ref["psi"]).max())
            rp, ci, hv = s.hessian(); rpo, cio, hvo, go = ora.hessian()
            print("   pattern equal", np.array_equal(rp, rpo) and np.array_equal(ci, cio), "nnzb", ci.size, "|dH|", np.abs(hv-hvo).max(), "Hmax", np.abs(hvo).max())
            lam = np.full((g.n_poses, 3), 1e-3); rhs = ref["gradient"]
            t=time.time(); w, it, rel = s.pcg_solve(lam, rhs); tp=time.time()-t
            wo = ora.linear_solve(lam, rhs)
            print("   pcg iters", it, "rel", rel, "time", tp, "|dw|/|w|", np.linalg.norm(w-wo)/np.linalg.norm(wo))
            t=time.time(); x, sm, tr = s.solve(); ts=time.time()-t
            fc = float(z[f"final_cost_dcs{dcs}"])
            print("   solve", ts, "s; final", sm.final_cost, "oracle", fc, "rel", abs(sm.final_cost-fc)/fc, "iters", sm.num_iterations, "ok", sm.num_successful_steps,
                  "pcg total", sm.total_pcg_iterations, sm.message.decode())
            okg = np.array([t_.step_is_successful for t_ in tr]); oko = z[f"trace_ok_dcs{dcs}"]
            print("   accept seq equal", np.array_equal(okg, oko), "max pose diff", np.abs(x - z[f"final_pose_dcs{dcs}"]).max())
            cg = np.array([t_.cost for t_ in tr]); co = z[f"trace_cost_dcs{dcs}"]
            n = min(len(cg), len(co)); print("   max rel trace cost diff", np.max(np.abs(cg[:n]-co[:n])/co[:n]))

# scale test
for N in (100_000, 1_000_000):
    t=time.time(); g = D.Graph.synthetic(N, int(2.7*N)+1, n_bogus=int(0.3*N)); print("gen", N, g.n_edges, time.time()-t)
    t=time.time(); s = D.Solver(g, dcs_on=True, max_num_iterations=3, pcg_max_iter=2000, pcg_rel_tol=1e-6); print("create", time.time()-t)
    s.linearize_resident(3)
    ms = s.linearize_resident(20)/20
    bytes_alg = 108*g.n_edges + 120*g.n_poses
    print(f"N={N} linearize {ms*1e3:.1f} us  edges/s {g.n_edges/ms*1e3:.3e}  alg GB/s {bytes_alg/ms/1e6:.1f}")
    t=time.time(); c, gr = s.linearize(g.pose_xyt); print("e2e linearize", time.time()-t, c)
    if N <= 100_000:
        ora = O.Oracle(g, dcs_on=True, num_threads=8); t=time.time(); ref = ora.evaluate(); print("oracle eval", time.time()-t, ref["cost"], abs(ref["cost"]-c)/c, np.abs(gr-ref["gradient"]).max())
    t=time.time(); x, sm, tr = s.solve(); print("solve 3 iters", time.time()-t, sm.initial_cost, sm.final_cost, "pcg", sm.total_pcg_iterations, "lin time", sm.linear_solver_time_s, "eval", sm.eval_time_s)
    if sm.total_pcg_iterations: print("  per pcg iter us", sm.linear_solver_time_s/sm.total_pcg_iterations*1e6)
    s.close()

"""k_pcg_cluster phase breakdown on a small graph: with the development library (DCS_B200_LIB=.../libdcs_b200_dev.so) the
kernel records the cycles thread 0 of CTA 0 spends in each phase of the iteration loop; with the product library the
script is just a short driver for ncu (one linearisation, two linear solves)."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
from dcs_b200 import capi

names = sys.argv[1:] or ["M3500_100_seed1", "INTEL_50_seed1"]
lib = capi.load_library()
probe = getattr(lib, "dcs_debug_cluster_cycles", None) if "dev" in os.path.basename(capi.lib_path()) else None
PH = ["barrier_p", "product", "pq_exchange", "update_stage", "substitution", "z_rz_exchange", "direction"]
for name in names:
    g = D.Graph.load_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    rng = np.random.default_rng(3)
    rhs = rng.normal(0, 1, (g.n_poses, 3)); lam = np.full((g.n_poses, 3), 1e-3)
    with D.Solver(g, dcs_on=True) as s:
        s.linearize(g.pose_xyt)
        s.pcg_solve(lam, rhs)
        t = time.perf_counter(); w, it, rel = s.pcg_solve(lam, rhs); dt = time.perf_counter() - t
        out = {"case": name, "n_poses": int(g.n_poses), "iterations": it, "rel_residual": rel, "us_per_iteration_wall": 1e6 * dt / max(1, it)}
        if probe is not None:
            c = (C.c_double * 8)()
            assert probe(c) == 0
            n = max(1.0, c[7])
            out["cycles_per_iteration"] = {PH[i]: round(c[i] / n, 1) for i in range(7)}
            out["cycles_per_iteration"]["total"] = round(sum(c[i] for i in range(7)) / n, 1)
        print(json.dumps(out), flush=True)

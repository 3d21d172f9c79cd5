"""For every libvar_*.so: bit-compare cost / gradient / Hessian blocks of a 20 K-pose graph with the default
library, then time k_linearize at 1 M poses.  One subprocess per library (DCS_B200_LIB)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, numpy as np
sys.path.insert(0, os.path.join(%r, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
g = D.Graph.synthetic(20000, 54001, n_bogus=6000)
with D.Solver(g, dcs_on=True) as s:
    cost, grad = s.linearize(g.pose_xyt); rp, ci, hv = s.hessian()
tag = os.path.basename(os.environ.get("DCS_B200_LIB", "default"))
np.savez("/tmp/vc_" + tag + ".npz", cost=cost, grad=grad, hv=hv)
N = 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
s = D.Solver(g, dcs_on=True)
s.linearize_resident(5)
print(tag, "linearize us %%.1f" %% (1e3 * s.linearize_resident(20) / 20), flush=True)
''' % ROOT
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "toy-robust-backend-slam_b200", "libvar_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib: env["DCS_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", code], env=env)
import numpy as np
ref = np.load("/tmp/vc_default.npz")
for lib in libs[1:]:
    v = np.load("/tmp/vc_" + os.path.basename(lib) + ".npz")
    print(os.path.basename(lib), "bit-identical:", all(np.array_equal(ref[k], v[k]) for k in ("cost", "grad", "hv")),
          "max |dH|", float(np.abs(ref["hv"] - v["hv"]).max()))

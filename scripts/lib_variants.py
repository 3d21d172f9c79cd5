import glob, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
N = 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
s = D.Solver(g, dcs_on=True)
s.linearize_resident(5)
print(os.path.basename(os.environ.get("DCS_B200_LIB", "default")), "linearize us %%.1f" %% (1e3 * s.linearize_resident(20) / 20), flush=True)
''' % ROOT
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "toy-robust-backend-slam_b200", "libvar_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib: env["DCS_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", code], env=env)

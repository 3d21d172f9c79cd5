import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
N = 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
s = D.Solver(g, dcs_on=True)
s.linearize_resident(5)
print("variant", os.environ.get("DCS_K1_VARIANT"), "linearize us", 1e3 * s.linearize_resident(20) / 20)
''' % ROOT
for v in (15, 31, 47, 63, 79, 127):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DCS_K1_VARIANT=str(v)))

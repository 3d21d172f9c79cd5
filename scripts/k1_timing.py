import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
os.environ["DCS_K1_DBG"] = os.path.join(ROOT, "gpurun_out", "k1dbg.bin")
import numpy as np, dcs_b200 as D
N = 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
s = D.Solver(g, dcs_on=True)
s.linearize_resident(5)
print("us", 1e3 * s.linearize_resident(20) / 20)
s.linearize_resident(1)
d = np.fromfile(os.environ["DCS_K1_DBG"], dtype=np.int64).reshape(-1, 8)
print("tasks", len(d), "prologue cyc mean", d[:,0].mean(), "loop", d[:,1].mean(), "write", d[:,2].mean(), "reduce", d[:,3].mean(), "deg0 mean", d[:,4].mean())
print("loop cycles per round (by deg):")
for k in range(2, 20):
    m = d[:,4] == k
    if m.sum() > 50: print(k, int(m.sum()), "loop", d[m,1].mean(), "per round", d[m,1].mean()/k, "prologue", d[m,0].mean(), "tail", (d[m,2]+d[m,3]).mean())
span = d[:,5].max() - d[:,5].min()
print("start span cycles", span, "total task cycles sum", d[:, :4].sum(), "per SM avg", d[:, :4].sum()/148)

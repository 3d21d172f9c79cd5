"""CPU prototype of the three-level additive preconditioner sketched in DESIGN.md section 8 (N4) - a study tool, not product.

  python scripts/precond_multilevel.py N_POSES RADIUS AGG1 AGG2 SWEEPS      e.g.  250000 1.38e5 64 16 4

  M^-1 = M_chain32^-1 + P1 ( S1 + P2 A2^-1 P2^T ) P1^T
  level 1: aggregates of AGG1 consecutive poses, modes x / y / theta (theta rotates the aggregate about its centroid),
           A1 = P1^T A P1, S1 = SWEEPS damped (0.7) 3x3-block-Jacobi sweeps from zero - a fixed polynomial, so M stays a
           fixed SPD operator and plain CG applies;
  level 2: aggregates of AGG2 level-1 nodes, same three modes, A2 = P2^T A1 P2 solved exactly (a few thousand unknowns at
           1 M poses: a dense inverse).
Prints the PCG iteration count to 1e-12 (compare with scripts/precond_study.py for block-Jacobi / chain-32 / exact two-level).
"""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from precond_study import D, O, hessian_csr, pcg


def prolong(agg, comp, cx, cy, x, y):
    nf, na = agg.shape[0], agg.max() + 1
    P = sp.csr_matrix((np.ones(nf), (np.arange(nf), agg * 3 + comp)), shape=(nf, 3 * na))
    mx, my = comp == 0, comp == 1
    r2 = np.concatenate([np.flatnonzero(mx), np.flatnonzero(my)]); c2 = np.concatenate([agg[mx] * 3 + 2, agg[my] * 3 + 2])
    v2 = np.concatenate([-(y[mx] - cy[agg[mx]]), x[my] - cx[agg[my]]])
    return P + sp.csr_matrix((v2, (r2, c2)), shape=(nf, 3 * na))


def main():
    n = int(sys.argv[1]); radius = float(sys.argv[2]); s1, s2, sweeps = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    k = n / 1_000_000
    g = D.Graph.synthetic(n, int(2_700_001 * k), n_bogus=int(300_000 * k))
    H, grad = hessian_csr(O.Oracle(g, dcs_on=True, num_threads=os.cpu_count() or 1))
    d = H.diagonal(); idx = np.flatnonzero(d > 0)
    A = (H + sp.diags(d / radius))[idx][:, idx].tocsr(); b = grad[idx]
    pose, comp, xyt = idx // 3, idx % 3, g.pose_xyt
    Ac = A.tocoo(); pr, pc = pose[Ac.row], pose[Ac.col]
    keep = (np.abs(pr - pc) <= 1) & (pr // 32 == pc // 32)
    Tlu = spla.splu(sp.csc_matrix((Ac.data[keep], (Ac.row[keep], Ac.col[keep])), shape=A.shape), permc_spec="NATURAL", diag_pivot_thresh=0.0)
    agg1 = pose // s1
    cnt = np.bincount(agg1); cx1 = np.bincount(agg1, weights=xyt[pose, 0]) / cnt; cy1 = np.bincount(agg1, weights=xyt[pose, 1]) / cnt
    P1 = prolong(agg1, comp, cx1, cy1, xyt[pose, 0], xyt[pose, 1])
    A1 = (P1.T @ A @ P1).tocsr()
    n1 = A1.shape[0]; node1, comp1 = np.arange(n1) // 3, np.arange(n1) % 3
    A1c = A1.tocoo(); kd = node1[A1c.row] == node1[A1c.col]
    D1 = spla.splu(sp.csc_matrix((A1c.data[kd], (A1c.row[kd], A1c.col[kd])), shape=A1.shape), permc_spec="NATURAL", diag_pivot_thresh=0.0)
    agg2 = node1 // s2
    cnt2 = np.bincount(agg2[::3]); cx2 = np.bincount(agg2[::3], weights=cx1) / cnt2; cy2 = np.bincount(agg2[::3], weights=cy1) / cnt2
    P2 = prolong(agg2, comp1, cx2, cy2, cx1[node1], cy1[node1])
    A2 = spla.splu((P2.T @ A1 @ P2).tocsc())
    print(f"level 1: {n1} unknowns, {A1.nnz} non-zeros; level 2: {P2.shape[1]} unknowns", flush=True)

    def coarse(r1):
        z = D1.solve(r1)
        for _ in range(sweeps - 1):
            z = z + 0.7 * D1.solve(r1 - A1 @ z)
        return z + P2 @ A2.solve(P2.T @ r1)

    t = time.time()
    _, it = pcg(A, b, lambda r: Tlu.solve(r) + P1 @ coarse(P1.T @ r))
    print(f"aggregates {s1} / {s2}, {sweeps} sweeps: {it} PCG iterations ({time.time() - t:.0f} s)")


if __name__ == "__main__":
    main()

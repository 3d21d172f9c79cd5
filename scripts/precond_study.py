"""CPU study of coarse-space preconditioners for the LM linear system (round 2; numbers in DESIGN.md section 8, N4).

  python scripts/precond_study.py N_POSES RADIUS AGG_SIZES [base]
  e.g.  python scripts/precond_study.py 250000 1.38e5 1024,256 base

Builds the benchmark graph family at N_POSES (same generator and ratios as bench.py), takes H = J^T J of the corrected
Jacobians at the initial point from the ORACLE (this is a study tool, not product code), forms A = H + diag(H) / RADIUS
(the LM system at the trust-region radius the 1 M-pose solve spends most of its iterations at) and counts PCG iterations
to 1e-12 for: 3x3 block-Jacobi, the chain-32 block-tridiagonal preconditioner of the product, and chain-32 plus an
additive aggregation coarse space (aggregates = AGG consecutive poses, piecewise-constant x / y / theta, optionally the
theta mode rotating the aggregate about its centroid), coarse system solved exactly (SuperLU).
"""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dcs_b200 as D, oracle_py as O


def hessian_csr(o):
    rp, ci, v, grad = o.hessian(None)
    N = rp.shape[0] - 1
    rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(rp))
    up = sp.bsr_matrix((v, ci, rp), shape=(3 * N, 3 * N)).tocsr()
    diag = rows == ci
    has = np.zeros(N, bool); has[rows[diag]] = True
    dg = sp.bsr_matrix((v[diag], np.flatnonzero(has), np.concatenate([[0], np.cumsum(has.astype(np.int64))])), shape=(3 * N, 3 * N)).tocsr()
    return (up + up.T - dg).tocsr(), grad.reshape(-1)


def pcg(A, b, Minv, tol=1e-12, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; b2 = np.sqrt(b @ b)
    for k in range(1, maxit + 1):
        q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.sqrt(r @ r) <= tol * b2: return x, k
        z = Minv(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit


def main():
    n = int(sys.argv[1]); radius = float(sys.argv[2]); aggs = [int(a) for a in sys.argv[3].split(",")]
    k = n / 1_000_000
    g = D.Graph.synthetic(n, int(2_700_001 * k), n_bogus=int(300_000 * k))
    H, grad = hessian_csr(O.Oracle(g, dcs_on=True, num_threads=os.cpu_count() or 1))
    d = H.diagonal(); idx = np.flatnonzero(d > 0)
    A = (H + sp.diags(d / radius))[idx][:, idx].tocsr(); b = grad[idx]
    pose, comp, nd = idx // 3, idx % 3, idx.shape[0]
    Ac = A.tocoo(); pr, pc = pose[Ac.row], pose[Ac.col]
    keep = (np.abs(pr - pc) <= 1) & (pr // 32 == pc // 32)
    Tlu = spla.splu(sp.csc_matrix((Ac.data[keep], (Ac.row[keep], Ac.col[keep])), shape=A.shape), permc_spec="NATURAL", diag_pivot_thresh=0.0)
    M1 = lambda r: Tlu.solve(r)
    if "base" in sys.argv:
        kj = pr == pc
        Jlu = spla.splu(sp.csc_matrix((Ac.data[kj], (Ac.row[kj], Ac.col[kj])), shape=A.shape), permc_spec="NATURAL", diag_pivot_thresh=0.0)
        print("block-Jacobi iterations", pcg(A, b, lambda r: Jlu.solve(r))[1], flush=True)
        print("chain-32 iterations", pcg(A, b, M1)[1], flush=True)
    xyt = g.pose_xyt
    for s in aggs:
        for rot in (0, 1):
            agg = pose // s; na = agg.max() + 1
            P = sp.csr_matrix((np.ones(nd), (np.arange(nd), agg * 3 + comp)), shape=(nd, 3 * na))
            if rot:
                cnt = np.bincount(agg)
                cx = np.bincount(agg, weights=xyt[pose, 0]) / cnt; cy = np.bincount(agg, weights=xyt[pose, 1]) / cnt
                mx, my = comp == 0, comp == 1
                r2 = np.concatenate([np.flatnonzero(mx), np.flatnonzero(my)]); c2 = np.concatenate([agg[mx] * 3 + 2, agg[my] * 3 + 2])
                v2 = np.concatenate([-(xyt[pose[mx], 1] - cy[agg[mx]]), xyt[pose[my], 0] - cx[agg[my]]])
                P = P + sp.csr_matrix((v2, (r2, c2)), shape=(nd, 3 * na))
            Acs = (P.T @ A @ P).tocsc()
            lu = spla.splu(Acs)
            t = time.time(); _, it = pcg(A, b, lambda r: M1(r) + P @ lu.solve(P.T @ r))
            print(f"aggregates of {s} poses, rotation mode {rot}: coarse unknowns {Acs.shape[0]}, fill {Acs.nnz / Acs.shape[0] ** 2:.3f}, "
                  f"iterations {it} ({time.time() - t:.0f} s)", flush=True)


if __name__ == "__main__":
    main()

"""CPU suite (-m "not gpu"): the oracle against the golden vectors, closed forms and finite differences."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_case
import oracle_py as O


@pytest.mark.parametrize("name", ["INTEL_50_seed1", "INTEL_0_seed1", "M3500_100_seed1"])
@pytest.mark.parametrize("dcs", [1, 0])
def test_jet_oracle_equals_reference_functors(name, dcs):
    """Golden r/J came from the REFERENCE's ceres_error.cpp compiled against oracle/ref_shim: the Jet restatement
    must reproduce them bit for bit (same operation order)."""
    g, z = load_case(name)
    ora = O.Oracle(g, dcs_on=bool(dcs))
    for tag, x in (("init", g.pose_xyt), ("pert", z["pose_perturbed"])):
        ev = ora.evaluate(x, raw=True)
        assert np.array_equal(ev["residuals"], z[f"ref_e_{tag}_dcs{dcs}"])
        assert np.array_equal(ev["jacobians"], z[f"ref_J_{tag}_dcs{dcs}"])


def test_live_reference_functors_if_built():
    """When oracle/_ref exists (build container, or shipped to the GPU box) compare live as well."""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    g, z = load_case("INTEL_50_seed1")
    kind = (g.kind != 0).astype(np.uint8)
    e, J = O.ref_edges(kind, g.meas_xyt, g.pose_xyt[g.edge_a], g.pose_xyt[g.edge_b])
    ev = O.Oracle(g, dcs_on=True).evaluate(raw=True)
    assert np.array_equal(e, ev["residuals"]) and np.array_equal(J, ev["jacobians"])


@pytest.mark.parametrize("dcs", [1, 0])
def test_closed_form_matches_jets(dcs):
    """Second derivation (analytic Jacobian incl. the derivative through psi) vs autodiff; tolerance widened
    by 1/cos^2(delta) on the folded angle row, edges with |cos delta| < 1e-4 skipped."""
    g, z = load_case("M3500_100_seed1")
    e_ref, J_ref = z[f"ref_e_pert_dcs{dcs}"], z[f"ref_J_pert_dcs{dcs}"]
    x = z["pose_perturbed"]
    worst = 0.0
    for k in range(0, g.n_edges, 3):
        a, b = g.edge_a[k], g.edge_b[k]
        use = bool(dcs) and g.kind[k] != 0
        e, J, _ = O.closed_form_edge(x[a], x[b], g.meas_xyt[k], use)
        cd = abs(np.cos(x[b, 2] - x[a, 2] - g.meas_xyt[k, 2]))
        if cd < 1e-4:
            continue
        tol = 1e-12 / cd ** 2
        assert np.all(np.abs(e - e_ref[k]) <= tol * np.maximum(1, np.abs(e_ref[k])))
        assert np.all(np.abs(J - J_ref[k]) <= tol * np.maximum(1, np.abs(J_ref[k])))
        worst = max(worst, np.abs(J - J_ref[k]).max())
    assert worst < 1e-11


def test_jacobian_vs_finite_differences():
    g, z = load_case("INTEL_50_seed1")
    ora = O.Oracle(g, dcs_on=True)
    x0 = z["pose_perturbed"].copy()
    J = ora.evaluate(x0, raw=True)["jacobians"]
    rng = np.random.default_rng(0)
    h = 1e-6
    for k in rng.choice(g.n_edges, 40, replace=False):
        a, b = g.edge_a[k], g.edge_b[k]
        if abs(np.cos(x0[b, 2] - x0[a, 2] - g.meas_xyt[k, 2])) < 1e-2:
            continue
        for col in range(6):
            xp, xm = x0.copy(), x0.copy()
            node, c = (a, col) if col < 3 else (b, col - 3)
            xp[node, c] += h; xm[node, c] -= h
            rp = ora.evaluate(xp, raw=True)["residuals"][k]; rm = ora.evaluate(xm, raw=True)["residuals"][k]
            fd = (rp - rm) / (2 * h)
            assert np.allclose(fd, J[k][:, col], rtol=1e-5, atol=1e-6), (k, col, fd, J[k][:, col])


def test_dcs_branch_and_huber_known_answers():
    # one edge, hand-computed: a at origin, b at (2,0,0), measurement identity -> e = (2,0,0)
    from dcs_b200 import Graph
    g = Graph(np.array([[0, 0, 0], [2.0, 0, 0]]), [0], [1], [[0, 0, 0]], [1])
    ev = O.Oracle(g, dcs_on=True).evaluate(raw=True)
    psi = np.sqrt(2 * 0.5 / (0.5 + 4.0))
    assert np.isclose(ev["psi"][0], psi, rtol=1e-15)
    assert np.allclose(ev["residuals"][0], [2 * psi, 0, 0], atol=1e-15)
    s = (2 * psi) ** 2
    assert np.isclose(ev["cost"], 0.5 * (2 * 0.01 * np.sqrt(s) - 1e-4), rtol=1e-14)
    assert np.isclose(ev["rho1"][0], 0.01 / np.sqrt(s), rtol=1e-14)
    # inlier: psi is the constant 1 (no derivative through psi), Huber quadratic region
    g2 = Graph(np.array([[0, 0, 0], [0.005, 0, 0]]), [0], [1], [[0, 0, 0]], [1])
    ev2 = O.Oracle(g2, dcs_on=True).evaluate()
    assert ev2["psi"][0] == 1.0 and ev2["rho1"][0] == 1.0
    assert np.isclose(ev2["cost"], 0.5 * 0.005 ** 2, rtol=1e-14)
    # odometry edges never get DCS, even with dcs_on
    g3 = Graph(np.array([[0, 0, 0], [2.0, 0, 0]]), [0], [1], [[0, 0, 0]], [0])
    assert O.Oracle(g3, dcs_on=True).evaluate()["psi"][0] == 1.0


def test_asin_fold():
    from dcs_b200 import Graph
    for d, want, sig in ((2.0, np.pi - 2.0, -1.0), (-2.5, -np.pi + 2.5, -1.0), (1.0, 1.0, 1.0)):
        g = Graph(np.array([[0, 0, 0], [0, 0, d]]), [0], [1], [[0, 0, 0]], [0])
        ev = O.Oracle(g, dcs_on=False).evaluate(raw=True)
        assert np.isclose(ev["residuals"][0, 2], want, atol=1e-14)
        assert np.isclose(ev["jacobians"][0, 2, 5], sig, atol=1e-12) and np.isclose(ev["jacobians"][0, 2, 2], -sig, atol=1e-12)


def test_structure_known_answers():
    """SURVEY §8d table (depends only on the data and the reader's |a-b|<5 rule)."""
    want = {"CSAIL": (1045, 1172, 1044, 128, 1171, 1170, 10), "FR079": (989, 1217, 988, 229, 1217, 1216, 10),
            "FRH": (1316, 2820, 2647, 173, 2820, 2817, 19), "INTEL": (1228, 1483, 1227, 256, 1483, 1482, 20),
            "M3500": (3500, 5453, 3609, 1844, 5453, 5450, 9), "MIT": (808, 827, 807, 20, 827, 826, 4)}
    st = json.load(open(os.path.join(GOLDEN, "structure.json")))
    for name, w in want.items():
        s = st[name]
        assert (s["n_poses"], s["n_edges"], s["n_odometry"], s["n_closure"], s["unique_pairs"], s["upper_offdiag"],
                s["max_degree"]) == w
        assert s["diag_blocks"] == s["n_poses"] - 1
        # oracle pattern over the committed edge lists
        z = np.load(os.path.join(GOLDEN, f"{name}_edges.npz"))
        from dcs_b200 import Graph
        E = z["edge_a"].shape[0]
        g = Graph(np.zeros((int(z["n_poses"]), 3)), z["edge_a"], z["edge_b"], np.zeros((E, 3)), np.zeros(E, np.uint8))
        rp, ci = O.Oracle(g).pattern()
        assert ci.size == s["nnzb"] == s["diag_blocks"] + s["upper_offdiag"]
        assert np.array_equal(rp, z["row_ptr"]) and np.array_equal(ci, z["col_idx"])


def test_hessian_is_jtj_and_linear_solve_exact():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    g, z = load_case("INTEL_50_seed1")
    ora = O.Oracle(g, dcs_on=True)
    rp, ci, hv, grad = ora.hessian()
    ev = ora.evaluate()
    # dense-ish J^T J from per-edge Jacobians
    N = g.n_poses
    rows, cols, vals = [], [], []
    for k in range(g.n_edges):
        for side, node in ((0, g.edge_a[k]), (1, g.edge_b[k])):
            for r in range(3):
                for c in range(3):
                    rows.append(3 * k + r); cols.append(3 * node + c); vals.append(ev["jacobians"][k, r, 3 * side + c])
    J = sp.csr_matrix((vals, (rows, cols)), shape=(3 * g.n_edges, 3 * N))
    H = (J.T @ J).toarray()
    for i in range(N):
        for q in range(rp[i], rp[i + 1]):
            j = ci[q]
            assert np.allclose(hv[q], H[3 * i:3 * i + 3, 3 * j:3 * j + 3], rtol=1e-11, atol=1e-12)
    assert np.allclose(grad[1:], (J.T @ ev["residuals"].ravel()).reshape(N, 3)[1:], rtol=1e-11, atol=1e-12)
    # exact solve vs scipy on the free block
    lam = np.full((N, 3), 1e-3)
    w = ora.linear_solve(lam, grad)
    free = np.arange(3, 3 * N)
    A = H[np.ix_(free, free)] + np.diag(lam.ravel()[free])
    w_ref = np.linalg.solve(A, grad.ravel()[free])
    assert np.allclose(w.ravel()[free], w_ref, rtol=1e-8, atol=1e-12)
    assert np.all(w[0] == 0)


@pytest.mark.parametrize("name,dcs", [("INTEL_50_seed1", 1), ("INTEL_50_seed1", 0), ("M3500_100_seed1", 1)])
def test_lm_trace_reproduces_golden(name, dcs):
    """The oracle's LM is deterministic: the committed trace must come back exactly (guards against drift of the
    checker itself).  Expected magnitudes also match SURVEY §6.2's independent scratch restatement."""
    g, z = load_case(name)
    x, s, tr = O.Oracle(g, dcs_on=bool(dcs)).solve()
    assert np.allclose([t.cost for t in tr], z[f"trace_cost_dcs{dcs}"], rtol=1e-12)
    assert np.array_equal([t.step_is_successful for t in tr], z[f"trace_ok_dcs{dcs}"])
    assert np.isclose(s.final_cost, float(z[f"final_cost_dcs{dcs}"]), rtol=1e-12)
    assert s.termination_type == 1 and s.num_iterations == 51


def test_survey_expectations():
    exp = {("INTEL_50_seed1", 1): (2.969102, 0.6679614), ("INTEL_50_seed1", 0): (39.64898, 2.940438),
           ("M3500_100_seed1", 1): (12.63069, 2.821174), ("M3500_100_seed1", 0): (81.31619, 9.154845),
           ("INTEL_0_seed1", 1): (2.471760, 0.1663066)}
    for (name, dcs), (c0, c1) in exp.items():
        _, z = load_case(name)
        assert np.isclose(float(z[f"cost_init_dcs{dcs}"]), c0, rtol=5e-7)
        assert np.isclose(float(z[f"final_cost_dcs{dcs}"]), c1, rtol=5e-7)


def test_multithreaded_oracle_matches_serial():
    g, _ = load_case("INTEL_50_seed1")
    a = O.Oracle(g, dcs_on=True, num_threads=1).hessian()
    b = O.Oracle(g, dcs_on=True, num_threads=4).hessian()
    assert np.allclose(a[2], b[2], rtol=1e-13, atol=1e-15) and np.allclose(a[3], b[3], rtol=1e-13, atol=1e-15)


def test_switchable_constraint_functors_match_reference():
    """METHOD 2 groundwork: the per-edge restatement of SwitchableClosureResidue / SwitchPriorResidue
    (reference src/ceres_error.cpp:199-317) is bit-identical to the reference's own compiled functors."""
    rng = np.random.default_rng(11)
    checked = 0
    for _ in range(400):
        pa = rng.normal(0, 20, 3); pb = pa + rng.normal(0, 2, 3); meas = rng.normal(0, 1, 3)
        pa[2] = rng.uniform(-3.1, 3.1); pb[2] = rng.uniform(-3.1, 3.1); meas[2] = rng.uniform(-1.5, 1.5)
        s = float(rng.uniform(-0.2, 1.3))
        e, J = O.sc_edge(pa, pb, meas, s)
        # structure: e = s * e_plain, d e / d s = e_plain, pose columns = s * plain Jacobian
        e0, J0 = O.sc_edge(pa, pb, meas, 1.0)
        assert np.allclose(e, s * e0, rtol=1e-15, atol=0) and np.array_equal(J[:, 6], e0)
        assert np.allclose(J[:, :6], s * J0[:, :6], rtol=1e-15, atol=0)
        ref = O.ref_sc_edge(pa, pb, meas, s)
        if ref is not None:
            assert np.array_equal(e, ref[0]) and np.array_equal(J, ref[1])
            checked += 1
    lam = 1.0
    for s in (1.0, 0.3, -0.1, 1.2):
        e, J = O.sc_prior(lam, s)
        assert e == np.sqrt(lam) * (1.0 - s) and J == -np.sqrt(lam)
        ref = O.ref_sc_prior(lam, s)
        if ref is not None:
            assert ref == (e, J)
    if O.ref_lib() is not None:
        assert checked == 400



def _dense_sc_lm(g, lam, max_it, huber=0.01):
    """Independent dense restatement of Ceres' default trust-region LM for METHOD 2: full Jacobian over
    (free poses, switches), Jacobi scaling, LM diagonal, dense solve.  Returns the per-iteration costs, the
    success flags, poses and switches."""
    N, E = g.n_poses, g.n_edges
    free = np.ones(N, bool); free[0] = False
    pidx = -np.ones(N, int); pidx[free] = np.arange(free.sum())
    loops = np.flatnonzero(g.kind != 0)
    sidx = {int(k): 3 * int(free.sum()) + i for i, k in enumerate(loops)}
    n = 3 * int(free.sum()) + len(loops)
    b2 = huber * huber

    def rho(sq):
        if sq > b2:
            r = np.sqrt(sq)
            return 2 * huber * r - b2, max(np.finfo(float).tiny, huber / r)
        return sq, 1.0

    def lin(x, sw):
        rows, r = [], []
        cost = 0.0
        for k in range(E):
            a, b = int(g.edge_a[k]), int(g.edge_b[k])
            s = sw[k] if g.kind[k] else 1.0
            e, J = O.sc_edge(x[a], x[b], g.meas_xyt[k], s)
            r0, r1 = rho(float(e @ e))
            cost += 0.5 * r0
            w = np.sqrt(r1)
            Jr = np.zeros((3, n))
            if free[a]: Jr[:, 3 * pidx[a]:3 * pidx[a] + 3] = w * J[:, 0:3]
            if free[b]: Jr[:, 3 * pidx[b]:3 * pidx[b] + 3] = w * J[:, 3:6]
            if g.kind[k]: Jr[:, sidx[k]] = w * J[:, 6]
            rows.append(Jr); r.append(w * e)
            if g.kind[k]:
                pr = np.sqrt(lam) * (1.0 - s)
                Jp = np.zeros((1, n)); Jp[0, sidx[k]] = -np.sqrt(lam)
                rows.append(Jp); r.append(np.array([pr])); cost += 0.5 * pr * pr
        return np.vstack(rows), np.concatenate(r), cost

    def cost_only(x, sw):
        c = 0.0
        for k in range(E):
            s = sw[k] if g.kind[k] else 1.0
            e, _ = O.sc_edge(x[int(g.edge_a[k])], x[int(g.edge_b[k])], g.meas_xyt[k], s)
            c += 0.5 * rho(float(e @ e))[0]
            if g.kind[k]: c += 0.5 * lam * (1.0 - s) ** 2
        return c

    def pack(x, sw): return np.concatenate([x[free].ravel(), sw[loops]])

    def unpack(v, x, sw):
        x = x.copy(); sw = sw.copy()
        x[free] = v[:3 * free.sum()].reshape(-1, 3); sw[loops] = v[3 * free.sum():]
        return x, sw

    x, sw = g.pose_xyt.copy(), np.ones(E)
    J, r, cost = lin(x, sw)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    radius, dec, reuse, diag = 1e4, 2.0, False, None
    costs, flags = [cost], []
    for it in range(max_it):
        Js = J * scale
        if not reuse: diag = np.clip((Js * Js).sum(0), 1e-6, 1e32)
        A = Js.T @ Js
        y = np.linalg.solve(A + np.diag(diag / radius), Js.T @ r)
        step = -y
        mcc = -(step @ (Js.T @ r)) - 0.5 * step @ A @ step
        reuse = True
        assert mcc > 0
        v = pack(x, sw) + step * scale
        xc, swc = unpack(v, x, sw)
        cand = cost_only(xc, swc)
        rel = (cost - cand) / mcc
        if rel > 1e-3:
            x, sw = xc, swc
            J, r, cost = lin(x, sw)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3)); dec = 2.0; reuse = False
            costs.append(cost); flags.append(1)
        else:
            radius /= dec; dec *= 2.0
            costs.append(cand); flags.append(0)
    return costs, flags, x, sw


def test_switchable_constraints_lm_elimination_matches_dense_full_system():
    """METHOD 2 groundwork: the oracle eliminates every switch inside the linear solve; an independent dense LM
    over the full (poses + switches) system must give the same iterates."""
    rng = np.random.default_rng(21)
    N = 40
    th = np.cumsum(rng.normal(0, 0.1, N)); xy = np.cumsum(np.c_[np.cos(th), np.sin(th)], axis=0)
    gt = np.c_[xy, th]
    ea = list(range(N - 1)); eb = list(range(1, N)); kind = [0] * (N - 1)
    for _ in range(25):
        a, b = sorted(rng.choice(N, 2, replace=False))
        if b - a > 5: ea.append(int(a)); eb.append(int(b)); kind.append(1 if rng.random() < 0.7 else 2)
    E = len(ea)
    meas = np.zeros((E, 3))
    for k in range(E):
        pa, pb = gt[ea[k]], gt[eb[k]]
        c, s = np.cos(pa[2]), np.sin(pa[2])
        d = pb[:2] - pa[:2]
        meas[k] = [c * d[0] + s * d[1], -s * d[0] + c * d[1], pb[2] - pa[2]]
        if kind[k] == 2: meas[k] = 0.0                          # bogus loops, like add_random_C
    meas += rng.normal(0, 0.01, meas.shape)
    pose = gt + rng.normal(0, 0.05, gt.shape); pose[0] = gt[0]
    import dcs_b200 as D
    g = D.Graph(pose, ea, eb, meas, kind)
    its = 12
    costs, flags, xd, swd = _dense_sc_lm(g, 1.0, its)
    ora = O.Oracle(g, dcs_on=False)
    x, sw, summ, trace = ora.sc_solve(lam=1.0, max_num_iterations=its, function_tolerance=0.0, parameter_tolerance=0.0,
                                      gradient_tolerance=0.0)
    got = [t.cost for t in trace]
    assert len(got) == its + 1
    assert [t.step_is_successful for t in trace[1:]] == flags
    assert np.allclose(got, costs, rtol=1e-9, atol=1e-14)
    loops = np.flatnonzero(np.array(kind) != 0)
    assert np.abs(sw[loops] - swd[loops]).max() < 1e-7 and np.abs(x - xd).max() < 1e-7
    # Huber(0.01) on the switched block leaves the switches a weak pull (rho' = 0.01 / |s e|): the bogus loops move
    # away from 1, the consistent ones stay there
    bog = np.flatnonzero(np.array(kind) == 2); good = np.flatnonzero(np.array(kind) == 1)
    if len(bog): assert sw[bog].mean() < sw[good].mean() - 0.05
    assert sw[good].min() > 0.9


@pytest.mark.parametrize("name", ["INTEL_50_seed1", "M3500_100_seed1"])
def test_method2_lm_trace_reproduces_golden(name):
    """METHOD 2 fixture for the GPU path to come (tests/golden/method2_traces.npz, made by make_golden.py method2):
    the oracle must reproduce it exactly."""
    z = np.load(os.path.join(GOLDEN, "method2_traces.npz"))
    g, _ = load_case(name)
    x, sw, s, tr = O.Oracle(g, dcs_on=False).sc_solve(lam=1.0)
    assert np.allclose([t.cost for t in tr], z[f"{name}_trace_cost"], rtol=1e-12)
    assert np.array_equal([t.step_is_successful for t in tr], z[f"{name}_trace_ok"])
    assert np.isclose(s.final_cost, float(z[f"{name}_final_cost"]), rtol=1e-12)
    loops = g.kind != 0
    assert np.allclose(sw[loops], z[f"{name}_switches"][loops], rtol=0, atol=1e-12)
    assert np.all(sw[~loops] == 1.0)                                     # odometry edges have no switch
    # METHOD 2 starts from the METHOD 0 cost (all switches 1) and ends below METHOD 0's final cost
    g0, z0 = load_case(name)
    assert np.isclose(tr[0].cost, float(z0["cost_init_dcs0"]), rtol=1e-12)
    assert s.final_cost < float(z0["final_cost_dcs0"])


@pytest.mark.parametrize("dcs", [0, 1])
def test_lm_matches_dense_restatement(dcs):
    """Second, independent restatement of the minimiser (dense numpy: full Jacobian, Jacobi scaling, LM diagonal,
    dense solve, Ceres' accept / radius rules) against the oracle's sparse-Cholesky LM, METHOD 0 and 1, on a graph
    small enough for dense algebra.  (Ceres itself is absent offline; this pins the oracle's linear algebra and
    bookkeeping to a second implementation, not to Ceres.)"""
    rng = np.random.default_rng(33 + dcs)
    N = 60
    th = np.cumsum(rng.normal(0, 0.15, N)); xy = np.cumsum(np.c_[np.cos(th), np.sin(th)], axis=0)
    gt = np.c_[xy, th]
    ea = list(range(N - 1)); eb = list(range(1, N)); kind = [0] * (N - 1)
    while len(ea) < N - 1 + 40:
        a, b = sorted(int(v) for v in rng.choice(N, 2, replace=False))
        if b - a > 5: ea.append(a); eb.append(b); kind.append(1 if rng.random() < 0.75 else 2)
    E = len(ea)
    meas = np.zeros((E, 3))
    for k in range(E):
        pa, pb = gt[ea[k]], gt[eb[k]]
        c, s = np.cos(pa[2]), np.sin(pa[2]); d = pb[:2] - pa[:2]
        meas[k] = [c * d[0] + s * d[1], -s * d[0] + c * d[1], pb[2] - pa[2]]
        if kind[k] == 2: meas[k] = 0.0
    meas += rng.normal(0, 0.02, meas.shape)
    pose = gt + rng.normal(0, 0.1, gt.shape); pose[0] = gt[0]
    import dcs_b200 as D
    g = D.Graph(pose, ea, eb, meas, kind)
    ora = O.Oracle(g, dcs_on=bool(dcs))
    free = np.ones(N, bool); free[0] = False
    col = -np.ones(N, int); col[free] = 3 * np.arange(free.sum())
    n = 3 * int(free.sum())

    def lin(x):
        ev = ora.evaluate(x)                          # corrected residuals / Jacobians per edge
        J = np.zeros((3 * E, n))
        for k in range(E):
            a, b = ea[k], eb[k]
            if free[a]: J[3 * k:3 * k + 3, col[a]:col[a] + 3] = ev["jacobians"][k][:, 0:3]
            if free[b]: J[3 * k:3 * k + 3, col[b]:col[b] + 3] = ev["jacobians"][k][:, 3:6]
        return J, ev["residuals"].ravel(), ev["cost"]

    its = 15
    x = g.pose_xyt.copy()
    J, r, cost = lin(x)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    radius, dec, reuse, diag = 1e4, 2.0, False, None
    costs, flags = [cost], []
    for _ in range(its):
        Js = J * scale
        if not reuse: diag = np.clip((Js * Js).sum(0), 1e-6, 1e32)
        A = Js.T @ Js
        step = -np.linalg.solve(A + np.diag(diag / radius), Js.T @ r)
        mcc = -(step @ (Js.T @ r)) - 0.5 * step @ A @ step
        reuse = True
        xc = x.copy(); xc[free] += (step * scale).reshape(-1, 3)
        cand = ora.cost(xc)
        rel = (cost - cand) / mcc
        if rel > 1e-3:
            x = xc; J, r, cost = lin(x)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3)); dec = 2.0; reuse = False
            costs.append(cost); flags.append(1)
        else:
            radius /= dec; dec *= 2.0
            costs.append(cand); flags.append(0)
    xo, s, tr = ora.solve(max_num_iterations=its, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
    assert [t.step_is_successful for t in tr[1:]] == flags
    assert np.allclose([t.cost for t in tr], costs, rtol=1e-9, atol=1e-14)
    assert sum(flags) >= 5


@pytest.mark.parametrize("fixture,seed", [("SYN10K_1000_s777", 777), ("SYN10K_1000", 12345)])
def test_syn10k_fixtures_are_the_generator_and_the_oracle(fixture, seed):
    """The 10 000-pose fixtures of the full-solve GPU test (tests/golden/make_golden.py syn10k): the committed graph is what
    the C++ generator + injector produce for that seed, the committed initial cost is the oracle's, and (one fixture, one
    iteration: the factorisation has 6.6 M non-zeros) the first LM step of the committed trace is reproduced."""
    import dcs_b200 as D
    z = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    g = D.Graph.synthetic(10000, 10688, n_bogus=1000, bogus_seed=seed)
    for k in ("pose_xyt", "edge_a", "edge_b", "meas_xyt", "kind"):
        assert np.array_equal(getattr(g, k), z[k]), k
    ora = O.Oracle(g, dcs_on=True, num_threads=os.cpu_count() or 1)
    assert np.isclose(ora.evaluate()["cost"], z["trace_cost_dcs1"][0], rtol=1e-13)
    if seed == 777:
        x, s, tr = ora.solve(max_num_iterations=1)
        assert np.allclose([t.cost for t in tr], z["trace_cost_dcs1"][:2], rtol=1e-12)
        assert tr[1].step_is_successful == z["trace_ok_dcs1"][1]
        assert int(s.factor_nnz) == int(z["factor_nnz"])

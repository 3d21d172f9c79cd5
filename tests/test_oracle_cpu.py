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


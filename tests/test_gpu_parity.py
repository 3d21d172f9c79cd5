"""GPU suite (-m gpu): the CUDA path, called through the C-ABI, against the oracle and the golden vectors.

Tolerances (fp64; SURVEY Appendix D, written out here):
  per-edge r, J, psi, rho'   |d| <= 1e-12 * max(1,|ref|) on the shipped datasets, widened by 1/cos^2(delta) on
                             everything that carries the folded angle; edges with |cos delta| < 1e-4 skipped
  DCS / Huber branch         exact, except edges within 1e-12 relative of the threshold
  block pattern              bit-exact
  H blocks, gradient         1e-11 relative to the largest entry
  LM trace                   same accept/reject sequence, iteration count and termination; cost 1e-9 relative
  final cost                 1e-9 relative (north_star), PCG tolerance 1e-12 relative residual
"""
import json
import os

import numpy as np
import pytest

import dcs_b200 as D
from dcs_b200 import Graph
import oracle_py as O
from conftest import GOLDEN, load_case

pytestmark = pytest.mark.gpu


def _cosd(g, x):
    return np.abs(np.cos(x[g.edge_b, 2] - x[g.edge_a, 2] - g.meas_xyt[:, 2]))


def _check_edges(g, x, ev, ref_r, ref_J, tol0=1e-12):
    cd = _cosd(g, x)
    keep = cd >= 1e-4
    tol = (tol0 / cd[keep] ** 2)
    dr = np.abs(ev["residuals"] - ref_r)[keep] / np.maximum(1, np.abs(ref_r[keep]))
    dJ = (np.abs(ev["jacobians"] - ref_J)[keep] / np.maximum(1, np.abs(ref_J[keep]))).reshape(keep.sum(), -1)
    assert (dr.max(axis=1) <= tol).all(), dr.max()
    assert (dJ.max(axis=1) <= tol).all(), dJ.max()


@pytest.mark.parametrize("name", ["INTEL_50_seed1", "INTEL_0_seed1", "M3500_100_seed1"])
@pytest.mark.parametrize("dcs", [1, 0])
def test_edge_residuals_jacobians_vs_oracle(name, dcs):
    g, z = load_case(name)
    ora = O.Oracle(g, dcs_on=bool(dcs))
    with D.Solver(g, dcs_on=bool(dcs)) as s:
        for x in (g.pose_xyt, z["pose_perturbed"]):
            ev = s.evaluate(x)
            ref = ora.evaluate(x)
            _check_edges(g, x, ev, ref["residuals"], ref["jacobians"])
            # branches: identical except within 1e-12 of the thresholds
            raw = ora.evaluate(x, raw=True)
            e2 = (raw["residuals"] ** 2).sum(axis=1)
            near_h = np.abs(e2 - 1e-4) <= 1e-12 * 1e-4
            assert np.array_equal((ev["rho1"] < 1)[~near_h], (ref["rho1"] < 1)[~near_h])
            assert np.array_equal(ev["psi"] < 1, ref["psi"] < 1)
            assert np.allclose(ev["psi"], ref["psi"], rtol=1e-13, atol=0)
            assert np.allclose(ev["rho1"], ref["rho1"], rtol=1e-11, atol=0)
            assert abs(ev["cost"] - ref["cost"]) <= 1e-13 * ref["cost"]
            gmax = np.abs(ref["gradient"]).max()
            assert np.abs(ev["gradient"] - ref["gradient"]).max() <= 1e-11 * gmax
            assert (ev["gradient"][g.fixed_pose] == 0).all()
            assert abs(s.cost(x) - ref["cost"]) <= 1e-13 * ref["cost"]          # cost-only kernel


@pytest.mark.parametrize("name", ["INTEL_50_seed1", "M3500_100_seed1"])
@pytest.mark.parametrize("dcs", [1, 0])
def test_raw_functor_values_vs_reference_golden(name, dcs):
    """Golden e/J are the REFERENCE's functors (pre-loss).  Undo the Huber scaling of the GPU output to compare."""
    g, z = load_case(name)
    with D.Solver(g, dcs_on=bool(dcs)) as s:
        for tag, x in (("init", g.pose_xyt), ("pert", z["pose_perturbed"])):
            ev = s.evaluate(x)
            sc = np.sqrt(ev["rho1"])
            raw = dict(residuals=ev["residuals"] / sc[:, None], jacobians=ev["jacobians"] / sc[:, None, None])
            _check_edges(g, x, raw, z[f"ref_e_{tag}_dcs{dcs}"], z[f"ref_J_{tag}_dcs{dcs}"], tol0=2e-12)


@pytest.mark.parametrize("name", ["CSAIL", "FR079", "FRH", "INTEL", "M3500", "MIT"])
def test_block_pattern_bit_exact(name):
    """CSAIL carries a duplicated pair, MIT has a > b edges, every set touches the constant pose 0."""
    z = np.load(os.path.join(GOLDEN, f"{name}_edges.npz"))
    st = json.load(open(os.path.join(GOLDEN, "structure.json")))[name]
    E = z["edge_a"].shape[0]
    g = Graph(np.zeros((int(z["n_poses"]), 3)), z["edge_a"], z["edge_b"], np.zeros((E, 3)), np.zeros(E, np.uint8))
    with D.Solver(g) as s:
        rp, ci = s.pattern()
    assert np.array_equal(rp, z["row_ptr"]) and np.array_equal(ci, z["col_idx"])
    assert ci.size == st["nnzb"] == st["diag_blocks"] + st["upper_offdiag"]
    rpo, cio = O.Oracle(g).pattern()
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)


@pytest.mark.parametrize("name,dcs", [("INTEL_50_seed1", 1), ("M3500_100_seed1", 1), ("M3500_100_seed1", 0)])
def test_hessian_and_gradient_vs_oracle(name, dcs):
    g, z = load_case(name)
    x = z["pose_perturbed"]
    with D.Solver(g, dcs_on=bool(dcs)) as s:
        s.linearize(x)
        rp, ci, hv = s.hessian()
    rpo, cio, hvo, go = O.Oracle(g, dcs_on=bool(dcs)).hessian(x)
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    assert np.abs(hv - hvo).max() <= 1e-11 * np.abs(hvo).max()
    # diagonal blocks are exactly symmetric
    diag = ci == np.repeat(np.arange(g.n_poses), np.diff(rp))
    assert np.array_equal(hv[diag], hv[diag].transpose(0, 2, 1))


def test_duplicate_edges_add_up_and_untouched_poses_are_ignored():
    # pose 4 is untouched; the pair (1,3) appears three times, once reversed; pose 0 constant
    pose = np.array([[0, 0, 0], [1, 0.1, 0.05], [2, -0.1, 0.1], [3, 0.2, -0.2], [9, 9, 9], [4, 0, 0.3]], float)
    ea = [0, 1, 2, 1, 3, 1, 3]; eb = [1, 2, 3, 3, 1, 3, 5]
    meas = np.array([[1, 0, 0], [1, 0, 0], [1, 0, 0], [2.1, 0, 0], [-2, 0.1, 0.0], [1.9, 0.2, 0.1], [1, 0, 0.4]], float)
    kind = [0, 0, 0, 1, 1, 2, 0]
    g = Graph(pose, ea, eb, meas, kind)
    ora = O.Oracle(g, dcs_on=True)
    with D.Solver(g, dcs_on=True) as s:
        ev = s.evaluate()
        rp, ci, hv = s.hessian()
        ref = ora.evaluate()
        rpo, cio, hvo, go = ora.hessian()
        assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
        assert rp[4] == rp[5] and rp[0] == rp[1]                 # untouched pose 4 and constant pose 0: no blocks
        assert np.allclose(hv, hvo, rtol=1e-12, atol=1e-14)
        assert np.allclose(ev["gradient"], ref["gradient"], rtol=1e-12, atol=1e-14)
        assert (ev["gradient"][4] == 0).all() and (ev["gradient"][0] == 0).all()
        x, sm, tr = s.solve()
        xo, so, to = ora.solve()
        assert np.array_equal(x[4], pose[4]) and np.array_equal(x[0], pose[0])
        assert sm.termination_type == so.termination_type and sm.num_iterations == so.num_iterations
        assert abs(sm.final_cost - so.final_cost) <= 1e-9 * max(so.final_cost, 1e-300) + 1e-18
        assert np.abs(x - xo).max() < 1e-6


def test_single_edge_and_empty_graph():
    g = Graph(np.array([[0, 0, 0], [1.2, 0.1, 0.2]]), [0], [1], [[1, 0, 0]], [1])
    with D.Solver(g, dcs_on=True) as s:
        ev = s.evaluate(); ref = O.Oracle(g, dcs_on=True).evaluate()
        assert np.allclose(ev["jacobians"], ref["jacobians"], rtol=1e-13, atol=1e-15)
        x, sm, _ = s.solve()
        assert sm.final_cost < 1e-15 and np.allclose(x[1], [1, 0, 0], atol=1e-7)
    g0 = Graph(np.zeros((3, 3)), [], [], np.zeros((0, 3)), [])
    with D.Solver(g0) as s:
        assert s.evaluate()["cost"] == 0.0
        rp, ci = s.pattern()
        assert ci.size == 0
        x, sm, _ = s.solve()
        assert sm.termination_type == 0 and sm.num_iterations == 1      # gradient tolerance at iteration 0


def test_hub_pose_with_many_loops():
    """One pose closes loops with 700 others: rows much longer than the CTA tile's average."""
    rng = np.random.default_rng(3)
    N = 1500
    th = np.cumsum(rng.normal(0, 0.05, N)); xy = np.cumsum(np.c_[np.cos(th), np.sin(th)], axis=0)
    pose = np.c_[xy, th]; pose[0] = 0
    ea = list(range(N - 1)); eb = list(range(1, N)); kind = [0] * (N - 1)
    hub = 700
    for j in range(0, N, 2):
        if abs(j - hub) > 5:
            ea.append(hub); eb.append(j); kind.append(1)
    E = len(ea)
    meas = rng.normal(0, 1, (E, 3))
    g = Graph(pose, ea, eb, meas, kind)
    with D.Solver(g, dcs_on=True) as s:
        ev = s.evaluate(); s.linearize(g.pose_xyt); rp, ci, hv = s.hessian()
    ora = O.Oracle(g, dcs_on=True)
    ref = ora.evaluate(); rpo, cio, hvo, _ = ora.hessian()
    assert np.array_equal(ci, cio)
    assert np.abs(hv - hvo).max() <= 1e-11 * np.abs(hvo).max()
    assert np.abs(ev["gradient"] - ref["gradient"]).max() <= 1e-11 * np.abs(ref["gradient"]).max()


@pytest.mark.parametrize("precond", [0, 1])
@pytest.mark.parametrize("name,dcs", [("INTEL_50_seed1", 1), ("M3500_100_seed1", 0)])
def test_pcg_matches_exact_cholesky(name, dcs, precond):
    """precond 0: 3x3 block-Jacobi; 1: block-Jacobi over 32-pose chain segments (block-tridiagonal, exact)."""
    g, z = load_case(name)
    ora = O.Oracle(g, dcs_on=bool(dcs))
    rhs = ora.evaluate()["gradient"]
    for lam_v in (1e-2, 1e-6):
        lam = np.full((g.n_poses, 3), lam_v)
        with D.Solver(g, dcs_on=bool(dcs), pcg_rel_tol=1e-13, preconditioner=precond) as s:
            s.linearize(g.pose_xyt)
            w, it, rel = s.pcg_solve(lam, rhs)
        wo = ora.linear_solve(lam, rhs)
        assert rel <= 1e-12 and it > 0
        assert np.linalg.norm(w - wo) <= 1e-9 * np.linalg.norm(wo)
        assert (w[g.fixed_pose] == 0).all()


@pytest.mark.parametrize("name,dcs", [("INTEL_50_seed1", 1), ("INTEL_50_seed1", 0), ("INTEL_0_seed1", 1),
                                      ("M3500_100_seed1", 1), ("M3500_100_seed1", 0)])
def test_full_lm_solve_matches_oracle_trace(name, dcs):
    """BASELINE.json configs 1 and 2 (real datasets + seeded outlier loops, DCS on vs off)."""
    g, z = load_case(name)
    with D.Solver(g, dcs_on=bool(dcs)) as s:
        x, sm, tr = s.solve()
    co = z[f"trace_cost_dcs{dcs}"]
    assert sm.num_iterations == len(co) == 51
    assert sm.termination_type == int(z[f"termination_dcs{dcs}"]) == 1
    assert np.array_equal([t.step_is_successful for t in tr], z[f"trace_ok_dcs{dcs}"])
    cg = np.array([t.cost for t in tr])
    assert (np.abs(cg - co) <= 1e-9 * co).all(), np.max(np.abs(cg - co) / co)
    rg = np.array([t.trust_region_radius for t in tr]); ro = z[f"trace_radius_dcs{dcs}"]
    assert np.allclose(rg, ro, rtol=1e-6)
    fc = float(z[f"final_cost_dcs{dcs}"])
    assert abs(sm.final_cost - fc) <= 1e-9 * fc
    assert np.abs(x - z[f"final_pose_dcs{dcs}"]).max() < 1e-6
    # inlier / outlier classification of loop edges at the final poses: psi < 1 (functor's own branch)
    if dcs:
        with D.Solver(g, dcs_on=True) as s2:
            psi = s2.evaluate(x)["psi"]
        ref_psi = z[f"final_psi_dcs{dcs}"]
        near = np.abs(ref_psi - 1.0) < 1e-6
        loops = g.kind != 0
        assert np.array_equal((psi < 1)[loops & ~near], (ref_psi < 1)[loops & ~near])
        bogus = g.kind == 2
        if bogus.any():
            assert (psi[bogus] < 0.5).mean() > 0.9          # DCS switches the injected loops off


def test_chain_preconditioner_cuts_iterations_and_keeps_the_trace():
    g, z = load_case("M3500_100_seed1")
    res = {}
    for pc in (0, 1):
        with D.Solver(g, dcs_on=True, preconditioner=pc, max_num_iterations=12) as s:
            x, sm, tr = s.solve()
        res[pc] = (sm.total_pcg_iterations, [t.step_is_successful for t in tr], sm.final_cost, x)
    assert res[0][1] == res[1][1] == list(z["trace_ok_dcs1"][:13])
    assert abs(res[0][2] - res[1][2]) <= 1e-10 * res[0][2]
    assert np.abs(res[0][3] - res[1][3]).max() < 1e-7
    assert res[1][0] < 0.75 * res[0][0], (res[0][0], res[1][0])


def test_solve_is_bit_reproducible():
    g, _ = load_case("INTEL_50_seed1")
    with D.Solver(g, dcs_on=True, max_num_iterations=8) as s:
        x1, s1, _ = s.solve()
    with D.Solver(g, dcs_on=True, max_num_iterations=8) as s:
        x2, s2, _ = s.solve()
    assert np.array_equal(x1, x2) and s1.final_cost == s2.final_cost


def test_synthetic_10k_stand_in_for_city10000():
    """BASELINE config 3: M10000.g2o is absent from the reference checkout (.MISSING_LARGE_BLOBS); a 10 000-pose
    synthetic graph with the public file's edge count (20 687) and 1000 outlier loops stands in."""
    g = Graph.synthetic(10000, 10688, n_bogus=1000)
    assert g.n_edges == 9999 + 10688 + 1000
    ora = O.Oracle(g, dcs_on=True)
    with D.Solver(g, dcs_on=True, max_num_iterations=10) as s:
        ev = s.evaluate(); ref = ora.evaluate()
        assert abs(ev["cost"] - ref["cost"]) <= 1e-12 * ref["cost"]
        x, sm, tr = s.solve()
    xo, so, to = ora.solve(max_num_iterations=10)
    assert np.array_equal([t.step_is_successful for t in tr], [t.step_is_successful for t in to])
    assert abs(sm.final_cost - so.final_cost) <= 1e-9 * so.final_cost


def test_full_size_1m_poses_4m_edges_properties():
    """BASELINE config 4 at full size: cost and gradient against the oracle, determinism, cost-only == cost,
    directional derivative."""
    N = 1_000_000
    g = Graph.synthetic(N, 2_700_001, n_bogus=300_000)
    assert g.n_edges == 4_000_000
    with D.Solver(g, dcs_on=True) as s:
        c1, g1 = s.linearize(g.pose_xyt)
        c2, g2 = s.linearize(g.pose_xyt)
        assert c1 == c2 and np.array_equal(g1, g2)                       # no atomics anywhere: bit-reproducible
        assert abs(s.cost(g.pose_xyt) - c1) <= 1e-12 * c1
        ref = O.Oracle(g, dcs_on=True).evaluate()
        assert abs(c1 - ref["cost"]) <= 1e-11 * ref["cost"]
        gm = np.abs(ref["gradient"]).max()
        assert np.abs(g1 - ref["gradient"]).max() <= 1e-9 * gm           # km-scale world: 1e-10 class (App. A.2)
        rng = np.random.default_rng(1)
        d = rng.normal(0, 1, g.pose_xyt.shape); d[0] = 0
        h = 1e-6
        fd = (s.cost(g.pose_xyt + h * d) - s.cost(g.pose_xyt - h * d)) / (2 * h)
        assert abs(fd - (g1 * d).sum()) <= 1e-5 * abs(fd)


def _write_g2o(path, g):
    """g2o text from a flat graph (information values are placeholders: METHOD 0/1 never read them)."""
    with open(path, "w") as f:
        for i, p in enumerate(g.pose_xyt):
            f.write("VERTEX_SE2 %d %.17g %.17g %.17g\n" % (i, p[0], p[1], p[2]))
        for a, b, m in zip(g.edge_a, g.edge_b, g.meas_xyt):
            f.write("EDGE_SE2 %d %d %.17g %.17g %.17g 1 0 0 1 0 1\n" % (a, b, m[0], m[1], m[2]))


def test_drop_in_cli_do_build(tmp_path):
    """`do_build.sh INTEL 50 1` (BASELINE config 1) end to end: same stdout lines, same four save/ files, final cost
    equal to the oracle's.  INTEL.g2o is rebuilt from the committed golden arrays (the reference checkout is not on
    the GPU box); DCS_SEED=1 reproduces the golden outlier loops (glibc rand())."""
    import subprocess
    from conftest import ROOT
    pkg = os.path.join(ROOT, "toy-robust-backend-slam_b200")
    g0, _ = load_case("INTEL_0_seed1")
    g50, z = load_case("INTEL_50_seed1")
    data = tmp_path / "data"; save = tmp_path / "save"
    data.mkdir(); save.mkdir()
    _write_g2o(str(data / "INTEL.g2o"), g0)
    env = dict(os.environ, DCS_SEED="1", DCS_DATA_PATH=str(data), DCS_SAVE_PATH=str(save))
    p = subprocess.run([os.path.join(pkg, "do_build.sh"), "INTEL", "50", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    out = p.stdout
    order = ["Start Reading PoseGraph", "Adding Bogus edges as described in Vertigo paper", "writePoseGraph nodes: ",
             "writePoseGraph Edges : ", "total nodes : 1228", "total nEdgesOdometry : 1227", "total nEdgesClosure : 256",
             "total nEdgesBogus : 50", "iter      cost      cost_change", "Termination:   NO_CONVERGENCE"]
    pos = [out.find(s) for s in order]
    assert all(q >= 0 for q in pos) and pos == sorted(pos), pos
    assert out.count("<--->") == 50
    a, b = g50.edge_a[-50:], g50.edge_b[-50:]
    assert "  %d<--->%d" % (a[0], b[0]) in out and "  %d<--->%d" % (a[-1], b[-1]) in out
    for f in ("init_nodes.txt", "init_edges.txt", "opt_nodes.txt", "opt_edges.txt"):
        assert (save / f).exists()
    init = np.genfromtxt(save / "init_nodes.txt"); opt = np.genfromtxt(save / "opt_nodes.txt")
    assert init.shape == opt.shape == (1228, 4)
    assert np.allclose(init[:, 1:], g50.pose_xyt, rtol=1e-5, atol=1e-5)                    # 6 significant digits
    assert np.allclose(opt[:, 1:], z["final_pose_dcs1"], rtol=2e-5, atol=2e-5)
    edges = np.genfromtxt(save / "opt_edges.txt", dtype=int)
    assert edges.shape == (1533, 3) and (edges[:, 2] == g50.kind).all()
    import re
    m = re.search(r"Final\s+([0-9.e+-]+)", out)
    assert m and abs(float(m.group(1)) - float(z["final_cost_dcs1"])) <= 1e-6 * float(z["final_cost_dcs1"])
    # METHOD 2/3/4 are refused, argc < 4 prints the usage
    p2 = subprocess.run([os.path.join(pkg, "build", "main"), "INTEL", "0", "3"], capture_output=True, text=True, env=env,
                        cwd=os.path.join(pkg, "build"))
    assert p2.returncode == 2

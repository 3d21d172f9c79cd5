"""CPU suite: the C++ host side (g2o reader, outlier injection, writers, synthetic generator) and the C-ABI
library's load/exports.  Reader / injector / writers are pinned against the REFERENCE's own header
(DCS-ceres/include/g2o_util.h compiled against oracle/ref_shim/boost) when the reference checkout is present."""
import ctypes as C
import filecmp
import os
import subprocess

import numpy as np
import pytest

import dcs_b200 as D
from dcs_b200 import Graph
from conftest import ROOT, load_case

REF_DATA = "/root/reference/DCS-ceres/data"
REF_READER = os.path.join(ROOT, "oracle", "_ref", "libdcs_ref_reader.so")
have_ref = os.path.isdir(REF_DATA) and os.path.exists(REF_READER)


def _ref_graph(path, seed, n_bogus):
    L = C.CDLL(REF_READER)
    L.ref_read.restype = C.c_void_p
    L.ref_read.argtypes = [C.c_char_p, C.c_uint, C.c_int]
    h = L.ref_read(path.encode(), seed, n_bogus)
    assert h
    n = (C.c_int * 4)()
    L.ref_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_counts(h, n)
    N, E = n[0], n[1] + n[2] + n[3]
    pose = np.empty((N, 3)); ea = np.empty(E, np.int32); eb = np.empty(E, np.int32); meas = np.empty((E, 3)); kind = np.empty(E, np.uint8)
    L.ref_flatten.argtypes = [C.c_void_p] * 6
    L.ref_flatten(h, *[a.ctypes.data_as(C.c_void_p) for a in (pose, ea, eb, meas, kind)])
    return L, h, tuple(n), pose, ea, eb, meas, kind


@pytest.mark.skipif(not have_ref, reason="reference checkout / oracle/_ref not present")
@pytest.mark.parametrize("name,n_bogus", [("INTEL", 50), ("M3500", 100), ("CSAIL", 7), ("FR079", 0), ("FRH", 3), ("MIT", 20)])
def test_reader_and_injection_match_reference(name, n_bogus, tmp_path):
    path = f"{REF_DATA}/{name}.g2o"
    L, h, counts, pose, ea, eb, meas, kind = _ref_graph(path, 1, n_bogus)
    g = Graph.from_g2o(path, n_bogus, seed=1)
    assert g.counts == counts
    assert np.array_equal(g.pose_xyt, pose) and np.array_equal(g.edge_a, ea) and np.array_equal(g.edge_b, eb)
    assert np.array_equal(g.meas_xyt, meas) and np.array_equal(g.kind, kind)
    # writers: byte-identical files for drawer/
    host = D.load_host_library()
    hh = host.dcs_host_read_g2o(path.encode())
    host.dcs_host_add_random_C(hh, n_bogus, 1, 1)
    mine_n, mine_e = str(tmp_path / "n.txt"), str(tmp_path / "e.txt")
    host.dcs_host_write_nodes(hh, mine_n.encode()); host.dcs_host_write_edges(hh, mine_e.encode())
    host.dcs_host_graph_free(hh)
    ref_n, ref_e = str(tmp_path / "rn.txt"), str(tmp_path / "re.txt")
    L.ref_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.ref_write(h, ref_n.encode(), ref_e.encode())
    assert filecmp.cmp(mine_n, ref_n, shallow=False) and filecmp.cmp(mine_e, ref_e, shallow=False)


@pytest.mark.skipif(not have_ref, reason="reference checkout not present")
def test_golden_graphs_are_what_the_reader_produces():
    g, z = load_case("INTEL_50_seed1")
    h = Graph.from_g2o(f"{REF_DATA}/INTEL.g2o", 50, seed=1)
    assert np.array_equal(g.edge_a, h.edge_a) and np.array_equal(g.meas_xyt, h.meas_xyt) and np.array_equal(g.pose_xyt, h.pose_xyt)
    assert tuple(z["counts"]) == (1228, 1227, 256, 50)   # DCS-ceres/docs/INTEL/info.txt:1-4 + 50 bogus


def test_parser_rules():
    text = ("VERTEX_SE2 0 0 0 0\nVERTEX2 1 1.5 -2 0.25\nVERTEX_SE2 2 3 4 5\n"
            "VERTEX_SE2  3   6 7 8\n"                       # runs of spaces compress
            "VERTEX_SE2 4 0 0 0\nVERTEX_SE2 5 0 0 0\nVERTEX_SE2 6 9 9 9\n"
            " VERTEX_SE2 7 1 1 1\n"                          # leading space -> empty first token -> ignored
            "FIX 0\n# comment\n"
            "EDGE_SE2 0 1 1 2 3 1 0 0 1 0 1\n"
            "EDGE2 6 1 -1 -2 -3 1 0 0 1 0 1\n"               # |a-b| = 5 -> closure
            "EDGE_SE2 5 1 0.5 0.5 0.5 1 0 0 1 0 1\n"         # |a-b| = 4 -> odometry, a > b kept as is
            "EDGE_SE2 2 3 7 8 9 1 0 0 1 0 1\n")
    g = Graph.from_g2o_text(text)
    assert g.counts == (7, 3, 1, 0)
    assert np.array_equal(g.pose_xyt[1], [1.5, -2, 0.25]) and np.array_equal(g.pose_xyt[3], [6, 7, 8])
    # order: odometry (file order) then closure
    assert g.edge_a.tolist() == [0, 5, 2, 6] and g.edge_b.tolist() == [1, 1, 3, 1]
    assert g.kind.tolist() == [0, 0, 0, 1]
    assert np.array_equal(g.meas_xyt[3], [-1, -2, -3])


def test_injection_is_glibc_rand_and_never_self_loop():
    text = "".join(f"VERTEX_SE2 {i} {i} 0 0\n" for i in range(7)) + "EDGE_SE2 0 1 1 0 0 1 0 0 1 0 1\n"
    g1 = Graph.from_g2o_text(text, n_bogus=500, seed=42)
    g2 = Graph.from_g2o_text(text, n_bogus=500, seed=42)
    g3 = Graph.from_g2o_text(text, n_bogus=500, seed=43)
    assert np.array_equal(g1.edge_a, g2.edge_a) and np.array_equal(g1.edge_b, g2.edge_b)
    assert not np.array_equal(g1.edge_a, g3.edge_a)
    assert g1.counts == (7, 1, 0, 500) and (g1.kind[1:] == 2).all()
    assert (g1.edge_a != g1.edge_b).all()
    assert (g1.meas_xyt[1:] == 0).all()          # rand()/RAND_MAX is an integer division
    # glibc: srand(42); rand() % 7 sequence, 5 draws per bogus edge
    libc = C.CDLL("libc.so.6")
    libc.srand(42)
    a = libc.rand() % 7; b = libc.rand() % 7
    if a == b:
        b = (b + 1) % 7
    assert (g1.edge_a[1], g1.edge_b[1]) == (a, b)
    for _ in range(3):
        libc.rand()
    a = libc.rand() % 7; b = libc.rand() % 7
    if a == b:
        b = (b + 1) % 7
    assert (g1.edge_a[2], g1.edge_b[2]) == (a, b)


def test_writer_format(tmp_path):
    text = "VERTEX_SE2 0 0.1234567891 -1e-7 3.14159265358979\nVERTEX_SE2 1 123456789 0 1\nEDGE_SE2 0 1 1 0 0 1 0 0 1 0 1\n"
    host = D.load_host_library()
    h = host.dcs_host_parse_g2o(text.encode(), len(text))
    host.dcs_host_write_nodes(h, str(tmp_path / "n.txt").encode())
    host.dcs_host_write_edges(h, str(tmp_path / "e.txt").encode())
    host.dcs_host_graph_free(h)
    # default ostream formatting = %g, 6 significant digits (reference g2o_util.h:98-101,184)
    assert open(tmp_path / "n.txt").read() == "0 0.123457 -1e-07 3.14159\n1 1.23457e+08 0 1\n"
    assert open(tmp_path / "e.txt").read() == "0 1 0\n"
    assert np.genfromtxt(tmp_path / "n.txt", usecols=(1, 2)).shape == (2, 2)   # what drawer/plot_results.py:28 does


@pytest.mark.skipif(not have_ref, reason="reference checkout / oracle/_ref not present")
def test_switches_writer_matches_reference(tmp_path):
    """METHOD 2's switches.txt (reference g2o_util.h:114-148): same bytes from the host writer."""
    path = f"{REF_DATA}/INTEL.g2o"
    L, h, counts, *_ = _ref_graph(path, 1, 20)
    n = counts[2] + counts[3]
    rng = np.random.default_rng(5)
    priors = np.ones(n)
    opt = np.clip(rng.normal(0.7, 0.4, n), -0.2, 1.3)
    opt[:3] = [1.0, 0.0, 1e-7]
    L.ref_write_switches.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
    ref_f, mine_f = str(tmp_path / "ref_sw.txt"), str(tmp_path / "sw.txt")
    L.ref_write_switches(h, ref_f.encode(), priors.ctypes.data_as(C.c_void_p), opt.ctypes.data_as(C.c_void_p), n)
    host = D.load_host_library()
    hh = host.dcs_host_read_g2o(path.encode())
    host.dcs_host_add_random_C(hh, 20, 1, 1)
    host.dcs_host_write_switches(hh, mine_f.encode(), priors.ctypes.data_as(C.c_void_p), opt.ctypes.data_as(C.c_void_p), n)
    host.dcs_host_graph_free(hh)
    assert filecmp.cmp(mine_f, ref_f, shallow=False)
    assert open(mine_f).readline() == "Odometry EDGES AHEAD\n"


def test_parallel_parser_keeps_serial_semantics(tmp_path):
    """The reader tokenises big files with all host threads; the result must be what a serial pass gives,
    including the acceptance rule (an edge is kept iff both ids are below the number of vertex lines read
    before it) across chunk boundaries.  Checked against a plain-Python serial pass and the reference header."""
    rng = np.random.default_rng(7)
    V, lines, nodes_seen, keep = 40000, [], 0, []
    vals = rng.normal(size=(V, 3)) * 100.0
    order = []          # interleave vertex and edge lines so every chunk holds both
    for i in range(V):
        lines.append(f"VERTEX_SE2 {i} {float(vals[i,0])!r} {vals[i,1]:.9g} {vals[i,2]:.17g}")
        nodes_seen += 1
        for _ in range(2):
            a, b = int(rng.integers(0, V)), int(rng.integers(0, V))     # about half of them point at later vertices
            m = [float(v) for v in rng.normal(size=3)]
            lines.append(f"EDGE_SE2 {a} {b} {m[0]!r} {m[1]!r} {m[2]!r} 44.721360 0 0 44.721360 0 44.721360")
            if a < nodes_seen and b < nodes_seen:
                keep.append((a, b, m[0], m[1], m[2]))
    text = "\n".join(lines) + "\n"
    assert len(text) > 8 << 20                                          # several parser threads
    g = Graph.from_g2o_text(text)
    assert g.n_poses == V and g.n_edges == len(keep)
    assert np.array_equal(g.pose_xyt[:, 0], vals[:, 0])                 # repr round-trips exactly
    odo = [k for k in keep if abs(k[0] - k[1]) < 5]; clo = [k for k in keep if abs(k[0] - k[1]) >= 5]
    want = np.array(odo + clo)
    assert np.array_equal(g.edge_a, want[:, 0].astype(np.int32)) and np.array_equal(g.edge_b, want[:, 1].astype(np.int32))
    assert np.array_equal(g.meas_xyt, want[:, 2:5])
    # the same file without the forward references (the reference indexes nNodes unchecked) through the
    # reference's own header, and through the mmap path
    ok = []
    seen = 0
    for ln in lines:
        if ln.startswith("VERTEX"):
            seen += 1; ok.append(ln)
        else:
            t = ln.split(" ")
            if int(t[1]) < seen and int(t[2]) < seen:
                ok.append(ln)
    path = str(tmp_path / "big.g2o")
    open(path, "w").write("\n".join(ok) + "\n")
    f = Graph.from_g2o(path)
    assert f.counts == g.counts and np.array_equal(f.meas_xyt, g.meas_xyt) and np.array_equal(f.pose_xyt, g.pose_xyt)
    assert np.array_equal(f.edge_a, g.edge_a) and np.array_equal(f.edge_b, g.edge_b)
    if have_ref:
        L, h, counts, pose, ea, eb, meas, kind = _ref_graph(path, 1, 0)
        assert g.counts == counts and np.array_equal(g.pose_xyt, pose) and np.array_equal(g.edge_a, ea)
        assert np.array_equal(g.edge_b, eb) and np.array_equal(g.meas_xyt, meas) and np.array_equal(g.kind, kind)


def test_synthetic_generator(tmp_path):
    g = Graph.synthetic(5000, 13501, n_bogus=1500)
    assert g.counts == (5000, 4999, 13501, 1500) and g.n_edges == 20000 and g.loops_made == 13501
    odo = g.kind == 0
    assert np.array_equal(g.edge_a[odo], np.arange(4999)) and np.array_equal(g.edge_b[odo], np.arange(1, 5000))
    loops = g.kind == 1
    assert (g.edge_b[loops] - g.edge_a[loops] > 5).all()          # (j, i) with j < i - 5
    # loop measurements are consistent with the ground truth up to the noise: relative translation <= sqrt(2)+noise
    assert np.linalg.norm(g.meas_xyt[loops, :2], axis=1).max() < 1.6
    # deterministic
    h = Graph.synthetic(5000, 13501, n_bogus=1500)
    assert np.array_equal(g.meas_xyt, h.meas_xyt) and np.array_equal(g.edge_a, h.edge_a)
    # round trip through the g2o text the reader consumes
    host = D.load_host_library()
    made = C.c_int64()
    hh = host.dcs_host_synth_manhattan(5000, 13501, 20260101, C.byref(made))
    path = str(tmp_path / "syn.g2o")
    assert host.dcs_host_write_g2o(hh, path.encode()) == 0
    host.dcs_host_graph_free(hh)
    r = Graph.from_g2o(path, 1500, seed=12345)
    assert r.counts == g.counts and np.array_equal(r.pose_xyt, g.pose_xyt) and np.array_equal(r.meas_xyt, g.meas_xyt)
    assert np.array_equal(r.edge_a, g.edge_a) and np.array_equal(r.kind, g.kind)


def test_cabi_library_loads_and_exports_every_declared_symbol():
    lib = D.load_library()
    for sym in D.DECLARED_SYMBOLS:
        assert hasattr(lib, sym), sym
    # header and binding list agree
    import re
    hdr = open(os.path.join(ROOT, "include", "dcs_b200.h")).read()
    declared = set(re.findall(r"\b(dcs_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(D.DECLARED_SYMBOLS), declared ^ set(D.DECLARED_SYMBOLS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", D.lib_path()]).decode()
    for sym in declared:
        assert f" T {sym}" in out
    assert b"sm_100a" in D.version().encode() or "sm_100a" in D.version()
    o = D.Options()
    lib.dcs_options_default(C.byref(o))
    assert (o.max_num_iterations, o.phi, o.huber_delta, o.initial_trust_region_radius) == (50, 0.5, 0.01, 1e4)
    assert o.function_tolerance == 1e-6 and o.gradient_tolerance == 1e-10 and o.parameter_tolerance == 1e-8


def _exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path]).decode()
    return {l.split()[2] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] in "TDB"}


def test_the_two_libraries_export_disjoint_symbols():
    """libdcs_b200.so exports exactly what include/dcs_b200.h declares (no dev probes), libdcs_host.so only its
    dcs_host_* entry points, and no name is defined by both (a process that links both must not mix
    cudaFreeHost with delete)."""
    cuda, host = _exported(D.lib_path()), _exported(D.host_lib_path())
    assert cuda == set(D.DECLARED_SYMBOLS), cuda ^ set(D.DECLARED_SYMBOLS)
    assert host and all(x.startswith("dcs_host_") for x in host), host
    assert not (cuda & host), cuda & host
    assert not any("debug" in x or "dbg" in x for x in cuda)


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must fail loudly, not compute on the CPU."""
    if D.device_count() > 0:
        pytest.skip("GPU present")
    g, _ = load_case("INTEL_50_seed1")
    with pytest.raises(D.DcsError) as ei:
        D.Solver(g)
    assert ei.value.code == 2     # DCS_ERR_CUDA


def test_bad_arguments_are_rejected_before_touching_cuda():
    g, _ = load_case("INTEL_50_seed1")
    bad = Graph(g.pose_xyt, g.edge_a.copy(), g.edge_b.copy(), g.meas_xyt, g.kind)
    bad.edge_b[5] = bad.edge_a[5]                      # self loop: Ceres would abort (reference g2o_util.h:160-163)
    with pytest.raises(D.DcsError) as ei:
        D.Solver(bad)
    assert ei.value.code == 1
    bad2 = Graph(g.pose_xyt, g.edge_a.copy(), g.edge_b.copy(), g.meas_xyt, g.kind)
    bad2.edge_a[0] = g.n_poses
    with pytest.raises(D.DcsError) as ei:
        D.Solver(bad2)
    assert ei.value.code == 1
    for fixed in (-2, g.n_poses):                       # -1 means "no constant pose"; anything else must be a pose
        with pytest.raises(D.DcsError) as ei:
            D.Solver(Graph(g.pose_xyt, g.edge_a, g.edge_b, g.meas_xyt, g.kind, fixed_pose=fixed))
        assert ei.value.code == 1
    with pytest.raises(D.DcsError) as ei:                # one NVSwitch node: at most 8 ranks per handle group
        D.Solver(g, rank=0, world=9, nccl_unique_id=b"\0" * 128)
    assert ei.value.code == 1


def test_product_does_not_reference_the_oracle():
    """The product tree must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "toy-robust-backend-slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "dcs_oracle" not in txt, os.path.join(dp, f)
    out = subprocess.check_output(["ldd", D.lib_path()]).decode()
    assert "oracle" not in out


def test_main_cli_usage_and_method_guard():
    exe = os.path.join(ROOT, "toy-robust-backend-slam_b200", "host", "main")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 255 and p.stdout.startswith("Usage: ")          # reference main.cpp:35-40 returns -1


def test_bench_trace_comparison_gates():
    """bench.py's N-rank vs 1-rank gate: leading iterations to 1e-9, count of agreeing iterations, worst deviation."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    mk = lambda c, ok: types.SimpleNamespace(cost=c, step_is_successful=ok)
    ref = [mk(10.0 / (i + 1), 1) for i in range(20)]
    same = [mk(t.cost * (1 + 1e-12), 1) for t in ref]
    ok, rel, k, agree, rel_all = bench.compare_traces(same, ref, 11)
    assert ok and k == 11 and agree == 20 and rel < 1e-11 and rel_all < 1e-11
    drift = [mk(t.cost * (1 + (1e-12 if i < 15 else 1e-6)), 1) for i, t in enumerate(ref)]
    ok, rel, k, agree, rel_all = bench.compare_traces(drift, ref, 11)
    assert ok and agree == 15 and 0.9e-6 < rel_all < 1.1e-6
    flip = [mk(t.cost, 0 if i == 5 else 1) for i, t in enumerate(ref)]
    ok, rel, k, agree, rel_all = bench.compare_traces(flip, ref, 11)
    assert not ok and agree == 5

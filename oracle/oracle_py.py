"""ctypes view of oracle/liboracle.so and oracle/_ref/libdcs_ref.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_DIR, "liboracle.so")
REF_LIB = os.path.join(_DIR, "_ref", "libdcs_ref.so")


class _Problem(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_edges", C.c_int32), ("pose_xyt", C.c_void_p), ("edge_a", C.c_void_p),
                ("edge_b", C.c_void_p), ("meas_xyt", C.c_void_p), ("kind", C.c_void_p), ("fixed_pose", C.c_int32),
                ("dcs_on", C.c_int32), ("phi", C.c_double), ("huber_delta", C.c_double), ("num_threads", C.c_int32)]


class LmOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("max_num_consecutive_invalid_steps", C.c_int32),
                ("jacobi_scaling", C.c_int32), ("verbose", C.c_int32)]


class OIteration(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
                ("pad", C.c_int32), ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double),
                ("gradient_norm", C.c_double), ("step_norm", C.c_double), ("relative_decrease", C.c_double),
                ("trust_region_radius", C.c_double), ("iteration_time_s", C.c_double), ("cumulative_time_s", C.c_double)]


class OSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("num_iterations", C.c_int32),
                ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
                ("termination_type", C.c_int32), ("total_time_s", C.c_double), ("eval_time_s", C.c_double),
                ("linear_solver_time_s", C.c_double), ("factor_nnz", C.c_int64), ("message", C.c_char * 128)]


def build():
    """Compiles liboracle.so (and _ref/libdcs_ref.so when the reference checkout is present)."""
    subprocess.check_call(["make", "-s", "-C", _DIR])
    if os.path.isdir("/root/reference/DCS-ceres/src"):
        subprocess.check_call(["make", "-s", "-C", _DIR, "ref"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_time_linearize.restype = C.c_double
        _lib.oracle_lm_options_default.restype = None
        _lib.oracle_edge_closed_form.restype = None
    return _lib


def ref_lib():
    """The reference's own functors (compiled from /root/reference against the ref_shim headers). None if absent."""
    global _ref
    if _ref is None and os.path.exists(REF_LIB):
        _ref = C.CDLL(REF_LIB)
    return _ref


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, graph, dcs_on=True, phi=0.5, huber_delta=0.01, num_threads=1):
        self.g = graph
        self.L = lib()
        self.p = _Problem(graph.n_poses, graph.n_edges, _ptr(graph.pose_xyt), _ptr(graph.edge_a), _ptr(graph.edge_b),
                          _ptr(graph.meas_xyt), _ptr(graph.kind), graph.fixed_pose, 1 if dcs_on else 0, phi,
                          huber_delta, num_threads)

    def evaluate(self, pose_xyt=None, raw=False):
        N, E = self.g.n_poses, self.g.n_edges
        x = None if pose_xyt is None else np.ascontiguousarray(pose_xyt, dtype=np.float64)
        cost = C.c_double()
        r = np.empty((E, 3)); J = np.empty((E, 3, 6)); psi = np.empty(E); rho1 = np.empty(E); g = np.empty((N, 3))
        rc = self.L.oracle_evaluate(C.byref(self.p), _ptr(x), 1 if raw else 0, C.byref(cost), _ptr(r), _ptr(J), _ptr(psi),
                                    _ptr(rho1), _ptr(g))
        assert rc == 0, rc
        return dict(cost=cost.value, residuals=r, jacobians=J, psi=psi, rho1=rho1, gradient=g)

    def cost(self, pose_xyt=None):
        x = None if pose_xyt is None else np.ascontiguousarray(pose_xyt, dtype=np.float64)
        c = C.c_double()
        assert self.L.oracle_cost(C.byref(self.p), _ptr(x), C.byref(c)) == 0
        return c.value

    def pattern(self):
        nnzb = C.c_int32()
        assert self.L.oracle_pattern(C.byref(self.p), C.byref(nnzb), None, None) == 0
        rp = np.empty(self.g.n_poses + 1, np.int32); ci = np.empty(nnzb.value, np.int32)
        assert self.L.oracle_pattern(C.byref(self.p), C.byref(nnzb), _ptr(rp), _ptr(ci)) == 0
        return rp, ci

    def hessian(self, pose_xyt=None):
        rp, ci = self.pattern()
        x = None if pose_xyt is None else np.ascontiguousarray(pose_xyt, dtype=np.float64)
        v = np.empty((ci.shape[0], 3, 3)); g = np.empty((self.g.n_poses, 3))
        assert self.L.oracle_hessian(C.byref(self.p), _ptr(x), _ptr(v), _ptr(g)) == 0
        return rp, ci, v, g

    def linear_solve(self, lam, rhs, pose_xyt=None):
        N = self.g.n_poses
        x = None if pose_xyt is None else np.ascontiguousarray(pose_xyt, dtype=np.float64)
        lam = None if lam is None else np.ascontiguousarray(lam, dtype=np.float64).reshape(N, 3)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64).reshape(N, 3)
        w = np.empty((N, 3))
        rc = self.L.oracle_linear_solve(C.byref(self.p), _ptr(x), _ptr(lam), _ptr(rhs), _ptr(w))
        assert rc == 0, rc
        return w

    def time_linearize(self, repeats=1):
        return float(self.L.oracle_time_linearize(C.byref(self.p), repeats))

    def solve(self, pose_xyt=None, verbose=False, **opts):
        x = np.array(self.g.pose_xyt if pose_xyt is None else pose_xyt, dtype=np.float64, order="C")
        o = LmOptions()
        self.L.oracle_lm_options_default(C.byref(o))
        o.verbose = 1 if verbose else 0
        for k, v in opts.items():
            setattr(o, k, v)
        s = OSummary()
        cap = o.max_num_iterations + 2
        trace = (OIteration * cap)()
        rc = self.L.oracle_solve(C.byref(self.p), C.byref(o), _ptr(x), C.byref(s), trace, cap)
        assert rc == 0, rc
        return x, s, [trace[i] for i in range(min(cap, s.num_iterations))]


def _sc_solve(self, switches=None, lam=1.0, pose_xyt=None, verbose=False, **opts):
    """METHOD 2: LM over poses + one switch per loop edge (oracle_sc_solve). Returns x, switches[E], summary, trace."""
    x = np.array(self.g.pose_xyt if pose_xyt is None else pose_xyt, dtype=np.float64, order="C")
    sw = np.ones(self.g.n_edges) if switches is None else np.array(switches, dtype=np.float64, order="C")
    o = LmOptions()
    self.L.oracle_lm_options_default(C.byref(o))
    o.verbose = 1 if verbose else 0
    for k, v in opts.items():
        setattr(o, k, v)
    s = OSummary()
    cap = o.max_num_iterations + 2
    trace = (OIteration * cap)()
    rc = self.L.oracle_sc_solve(C.byref(self.p), C.c_double(lam), C.byref(o), _ptr(x), _ptr(sw), C.byref(s), trace, cap)
    assert rc == 0, rc
    return x, sw, s, [trace[i] for i in range(min(cap, s.num_iterations))]


Oracle.sc_solve = _sc_solve


def closed_form_edge(pa, pb, meas, dcs, phi=0.5):
    pa = np.ascontiguousarray(pa, dtype=np.float64); pb = np.ascontiguousarray(pb, dtype=np.float64)
    meas = np.ascontiguousarray(meas, dtype=np.float64)
    e = np.empty(3); J = np.empty((3, 6)); psi = C.c_double()
    lib().oracle_edge_closed_form(_ptr(pa), _ptr(pb), _ptr(meas), C.c_int(1 if dcs else 0), C.c_double(phi), _ptr(e), _ptr(J),
                                  C.byref(psi))
    return e, J, psi.value


def ref_edges(kind, meas, pa, pb, jac=True):
    """Evaluates the REFERENCE's compiled functors. kind[k]: 0 OdometryResidue, 1 DCSClosureResidue."""
    R = ref_lib()
    if R is None:
        return None
    kind = np.ascontiguousarray(kind, dtype=np.uint8)
    meas = np.ascontiguousarray(meas, dtype=np.float64); pa = np.ascontiguousarray(pa, dtype=np.float64)
    pb = np.ascontiguousarray(pb, dtype=np.float64)
    n = kind.shape[0]
    e = np.empty((n, 3)); J = np.empty((n, 3, 6)) if jac else None
    rc = R.ref_edges(C.c_int(n), _ptr(kind), _ptr(meas), _ptr(pa), _ptr(pb), _ptr(e), _ptr(J))
    assert rc == 0
    return e, J


def sc_edge(pa, pb, meas, s):
    """METHOD 2 per-edge restatement (oracle_sc_edge): e[3], J[3,7] (P1, P2, s), no loss."""
    pa = np.ascontiguousarray(pa, dtype=np.float64); pb = np.ascontiguousarray(pb, dtype=np.float64)
    meas = np.ascontiguousarray(meas, dtype=np.float64)
    e = np.empty(3); J = np.empty((3, 7))
    L = lib()
    L.oracle_sc_edge.restype = None
    L.oracle_sc_edge(_ptr(pa), _ptr(pb), _ptr(meas), C.c_double(s), _ptr(e), _ptr(J))
    return e, J


def sc_prior(lam, s):
    e = np.empty(1); J = np.empty(1)
    L = lib()
    L.oracle_sc_prior.restype = None
    L.oracle_sc_prior(C.c_double(lam), C.c_double(s), _ptr(e), _ptr(J))
    return e[0], J[0]


def ref_sc_edge(pa, pb, meas, s):
    """The REFERENCE's compiled SwitchableClosureResidue through the shim (None when oracle/_ref is absent)."""
    R = ref_lib()
    if R is None:
        return None
    pa = np.ascontiguousarray(pa, dtype=np.float64); pb = np.ascontiguousarray(pb, dtype=np.float64)
    meas = np.ascontiguousarray(meas, dtype=np.float64)
    e = np.empty(3); J = np.empty((3, 7))
    assert R.ref_switchable_edge(_ptr(meas), _ptr(pa), _ptr(pb), C.c_double(s), _ptr(e), _ptr(J)) == 0
    return e, J


def ref_sc_prior(lam, s):
    R = ref_lib()
    if R is None:
        return None
    e = np.empty(1); J = np.empty(1)
    assert R.ref_switch_prior(C.c_double(lam), C.c_double(s), _ptr(e), _ptr(J)) == 0
    return e[0], J[0]

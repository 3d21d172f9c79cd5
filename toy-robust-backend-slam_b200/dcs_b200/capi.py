"""ctypes bindings: libdcs_b200.so (include/dcs_b200.h) and libdcs_host.so (host/host_capi.cpp)."""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lib_path():
    # DCS_B200_LIB: development override (kernel variants); the default is the in-tree build
    return os.environ.get("DCS_B200_LIB") or os.path.join(_PKG, "libdcs_b200.so")


def host_lib_path():
    return os.path.join(_PKG, "host", "libdcs_host.so")


class DcsError(RuntimeError):
    def __init__(self, code, where, text):
        super().__init__(f"{where} failed with status {code}: {text}")
        self.code = code


# ---- C structs (field order mirrors include/dcs_b200.h) ---------------------------------------------
class _Graph(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_edges", C.c_int32), ("pose_xyt", C.c_void_p), ("edge_a", C.c_void_p),
                ("edge_b", C.c_void_p), ("meas_xyt", C.c_void_p), ("kind", C.c_void_p), ("fixed_pose", C.c_int32)]


class Options(C.Structure):
    _fields_ = [("dcs_on", C.c_int32), ("phi", C.c_double), ("huber_delta", C.c_double),
                ("max_num_iterations", C.c_int32), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
                ("pcg_rel_tol", C.c_double), ("pcg_max_iter", C.c_int32), ("pcg_check_every", C.c_int32),
                ("preconditioner", C.c_int32),
                ("device", C.c_int32), ("verbose", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("max_solver_time_s", C.c_double),
                ("switchable_on", C.c_int32), ("switch_prior_lambda", C.c_double)]


class Iteration(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
                ("linear_solver_iterations", C.c_int32), ("cost", C.c_double), ("cost_change", C.c_double),
                ("gradient_max_norm", C.c_double), ("gradient_norm", C.c_double), ("step_norm", C.c_double),
                ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double),
                ("linear_solver_residual", C.c_double), ("iteration_time_s", C.c_double),
                ("cumulative_time_s", C.c_double), ("linear_solver_true_residual", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("num_iterations", C.c_int32),
                ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
                ("termination_type", C.c_int32), ("total_pcg_iterations", C.c_int64), ("total_time_s", C.c_double),
                ("eval_time_s", C.c_double), ("linear_solver_time_s", C.c_double), ("message", C.c_char * 128)]


class BatchItem(C.Structure):
    _fields_ = [("graph", _Graph), ("pose_xyt_inout", C.c_void_p), ("summary", Summary), ("status", C.c_int32)]


# every symbol include/dcs_b200.h declares (tests check the .so exports all of them)
DECLARED_SYMBOLS = ["dcs_options_default", "dcs_version", "dcs_device_count", "dcs_partition", "dcs_nccl_unique_id", "dcs_create",
                    "dcs_destroy", "dcs_evaluate", "dcs_linearize", "dcs_linearize_resident", "dcs_cost",
                    "dcs_get_pattern", "dcs_get_hessian", "dcs_pcg_solve", "dcs_solve", "dcs_solve_batch", "dcs_get_switches", "dcs_host_alloc", "dcs_host_free",
                    "dcs_last_error",
                    "dcs_launch_count"]

_lib = None
_host = None


def load_library():
    """Loads libdcs_b200.so. Raises if it was not built — there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
                          "There is no CPU fallback for the DCS-LM path.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.dcs_version.restype = C.c_char_p
    lib.dcs_last_error.restype = C.c_char_p
    lib.dcs_launch_count.restype = C.c_int64
    lib.dcs_launch_count.argtypes = [C.c_int]
    lib.dcs_options_default.argtypes = [C.POINTER(Options)]
    lib.dcs_options_default.restype = None
    lib.dcs_create.argtypes = [C.POINTER(_Graph), C.POINTER(Options), C.POINTER(C.c_void_p)]
    lib.dcs_destroy.argtypes = [C.c_void_p]
    lib.dcs_destroy.restype = None
    vp = C.c_void_p
    lib.dcs_evaluate.argtypes = [vp, vp, C.POINTER(C.c_double), vp, vp, vp, vp, vp]
    lib.dcs_linearize.argtypes = [vp, vp, C.POINTER(C.c_double), vp]
    lib.dcs_linearize_resident.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
    lib.dcs_cost.argtypes = [vp, vp, C.POINTER(C.c_double)]
    lib.dcs_get_pattern.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, vp]
    lib.dcs_get_hessian.argtypes = [vp, vp]
    lib.dcs_pcg_solve.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    lib.dcs_solve.argtypes = [vp, vp, C.POINTER(Summary), C.POINTER(Iteration), C.c_int32]
    lib.dcs_get_switches.argtypes = [vp, vp]
    lib.dcs_solve_batch.argtypes = [C.POINTER(BatchItem), C.c_int32, C.POINTER(Options), C.c_int32]
    lib.dcs_nccl_unique_id.argtypes = [vp]
    lib.dcs_host_alloc.restype = C.c_void_p
    lib.dcs_host_alloc.argtypes = [C.c_uint64]
    lib.dcs_host_free.argtypes = [C.c_void_p]
    lib.dcs_host_free.restype = None
    _lib = lib
    return lib


def load_host_library():
    global _host
    if _host is not None:
        return _host
    path = host_lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with __graft_entry__.build()")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.dcs_host_read_g2o.restype = vp
    lib.dcs_host_read_g2o.argtypes = [C.c_char_p]
    lib.dcs_host_parse_g2o.restype = vp
    lib.dcs_host_parse_g2o.argtypes = [C.c_char_p, C.c_int64]
    lib.dcs_host_synth_manhattan.restype = vp
    lib.dcs_host_synth_manhattan.argtypes = [C.c_int32, C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    lib.dcs_host_add_random_C.argtypes = [vp, C.c_int32, C.c_uint32, C.c_int32]
    lib.dcs_host_add_random_C.restype = None
    lib.dcs_host_counts.argtypes = [vp] + [C.POINTER(C.c_int32)] * 4
    lib.dcs_host_counts.restype = None
    lib.dcs_host_flatten.argtypes = [vp] * 6
    lib.dcs_host_flatten.restype = None
    lib.dcs_host_set_poses.argtypes = [vp, vp]
    lib.dcs_host_set_poses.restype = None
    lib.dcs_host_write_nodes.argtypes = [vp, C.c_char_p]
    lib.dcs_host_write_edges.argtypes = [vp, C.c_char_p]
    lib.dcs_host_write_switches.argtypes = [vp, C.c_char_p, vp, vp, C.c_int32]
    lib.dcs_host_write_switches.restype = None
    lib.dcs_host_write_g2o.argtypes = [vp, C.c_char_p]
    lib.dcs_host_graph_free.argtypes = [vp]
    lib.dcs_host_graph_free.restype = None
    _host = lib
    return lib


def version():
    return load_library().dcs_version().decode()


def device_count():
    return int(load_library().dcs_device_count())


def launch_count(reset=False):
    return int(load_library().dcs_launch_count(1 if reset else 0))


def partition(n_poses, n_edges, rank, world):
    """(row_lo, n_rows, rows_per_rank, edge_lo, edge_hi) of a rank; pure host arithmetic."""
    out = (C.c_int32 * 5)()
    lib = load_library()
    rc = lib.dcs_partition(n_poses, n_edges, rank, world, out)
    if rc:
        raise DcsError(rc, "dcs_partition", "bad argument")
    return tuple(out)


def pinned_empty(shape, dtype=np.float64):
    """numpy array over page-locked host memory from dcs_host_alloc (kept alive by the array's base object)."""
    lib = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib.dcs_host_alloc(max(n, 8))
    if not ptr:
        raise DcsError(2, "dcs_host_alloc", lib.dcs_last_error().decode())

    class _Owner:
        def __init__(self, p): self.p = p
        def __del__(self):
            try: lib.dcs_host_free(self.p)
            except Exception: pass

    buf = (C.c_char * n).from_address(ptr)
    buf._owner = _Owner(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    lib = load_library()
    rc = lib.dcs_nccl_unique_id(buf)
    if rc:
        raise DcsError(rc, "dcs_nccl_unique_id", lib.dcs_last_error().decode())
    return buf.raw


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Graph:
    """Flat pose graph in the reference's residual-block order (odometry, closure, bogus)."""

    def __init__(self, pose_xyt, edge_a, edge_b, meas_xyt, kind, fixed_pose=0):
        self.pose_xyt = np.ascontiguousarray(pose_xyt, dtype=np.float64).reshape(-1, 3)
        self.edge_a = np.ascontiguousarray(edge_a, dtype=np.int32)
        self.edge_b = np.ascontiguousarray(edge_b, dtype=np.int32)
        self.meas_xyt = np.ascontiguousarray(meas_xyt, dtype=np.float64).reshape(-1, 3)
        self.kind = np.ascontiguousarray(kind, dtype=np.uint8)
        self.fixed_pose = int(fixed_pose)

    n_poses = property(lambda s: s.pose_xyt.shape[0])
    n_edges = property(lambda s: s.edge_a.shape[0])

    # -- constructors over the C++ host code ---------------------------------------------------------
    @classmethod
    def _from_host(cls, host, h):
        n = [C.c_int32() for _ in range(4)]
        host.dcs_host_counts(h, *[C.byref(x) for x in n])
        N, E = n[0].value, n[1].value + n[2].value + n[3].value
        pose = np.empty((N, 3)); ea = np.empty(E, np.int32); eb = np.empty(E, np.int32)
        meas = np.empty((E, 3)); kind = np.empty(E, np.uint8)
        host.dcs_host_flatten(h, _ptr(pose), _ptr(ea), _ptr(eb), _ptr(meas), _ptr(kind))
        g = cls(pose, ea, eb, meas, kind, 0)
        g.counts = tuple(x.value for x in n)
        return g

    @classmethod
    def from_g2o(cls, path, n_bogus=0, seed=1):
        """host/g2o_util.h reader + srand(seed) + add_random_C(n_bogus)."""
        host = load_host_library()
        h = host.dcs_host_read_g2o(os.fsencode(path))
        if not h:
            raise FileNotFoundError(path)
        try:
            if n_bogus:
                host.dcs_host_add_random_C(h, n_bogus, seed, 1)
            return cls._from_host(host, h)
        finally:
            host.dcs_host_graph_free(h)

    @classmethod
    def from_g2o_text(cls, text, n_bogus=0, seed=1):
        host = load_host_library()
        data = text.encode() if isinstance(text, str) else text
        h = host.dcs_host_parse_g2o(data, len(data))
        try:
            if n_bogus:
                host.dcs_host_add_random_C(h, n_bogus, seed, 1)
            return cls._from_host(host, h)
        finally:
            host.dcs_host_graph_free(h)

    @classmethod
    def synthetic(cls, n_poses, n_loops, n_bogus=0, gen_seed=20260101, bogus_seed=12345):
        """host/synth.h Manhattan world + the same outlier injection."""
        host = load_host_library()
        made = C.c_int64()
        h = host.dcs_host_synth_manhattan(n_poses, n_loops, gen_seed, C.byref(made))
        if not h:
            raise ValueError("synthetic generator failed")
        try:
            if n_bogus:
                host.dcs_host_add_random_C(h, n_bogus, bogus_seed, 1)
            g = cls._from_host(host, h)
            g.loops_made = made.value
            return g
        finally:
            host.dcs_host_graph_free(h)

    def save_npz(self, path, **extra):
        np.savez_compressed(path, pose_xyt=self.pose_xyt, edge_a=self.edge_a, edge_b=self.edge_b,
                            meas_xyt=self.meas_xyt, kind=self.kind, fixed_pose=self.fixed_pose, **extra)

    @classmethod
    def load_npz(cls, path):
        z = np.load(path)
        return cls(z["pose_xyt"], z["edge_a"], z["edge_b"], z["meas_xyt"], z["kind"], int(z["fixed_pose"]))


class Solver:
    """dcs_create / dcs_evaluate / dcs_linearize / dcs_solve / dcs_destroy."""

    def __init__(self, graph, dcs_on=True, **opts):
        self.lib = load_library()
        self.graph = graph
        o = Options()
        self.lib.dcs_options_default(C.byref(o))
        o.dcs_on = 1 if dcs_on else 0
        self._uid = None
        for k, v in opts.items():
            if k == "nccl_unique_id":
                self._uid = C.create_string_buffer(v, 128)
                o.nccl_unique_id = C.cast(self._uid, C.c_void_p)
            else:
                setattr(o, k, v)
        self.options = o
        g = _Graph(graph.n_poses, graph.n_edges, _ptr(graph.pose_xyt), _ptr(graph.edge_a), _ptr(graph.edge_b),
                   _ptr(graph.meas_xyt), _ptr(graph.kind), graph.fixed_pose)
        self.h = C.c_void_p()
        self._ck(self.lib.dcs_create(C.byref(g), C.byref(o), C.byref(self.h)), "dcs_create")

    def _ck(self, rc, where):
        if rc:
            raise DcsError(rc, where, self.lib.dcs_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.dcs_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def evaluate(self, pose_xyt=None, residuals=True, jacobians=True, gradient=True):
        N, E = self.graph.n_poses, self.graph.n_edges
        x = None if pose_xyt is None else np.ascontiguousarray(pose_xyt, dtype=np.float64)
        cost = C.c_double()
        r = np.empty((E, 3)) if residuals else None
        J = np.empty((E, 3, 6)) if jacobians else None
        per_edge = residuals or jacobians
        psi = np.empty(E) if per_edge else None
        rho1 = np.empty(E) if per_edge else None
        g = np.empty((N, 3)) if gradient else None
        self._ck(self.lib.dcs_evaluate(self.h, _ptr(x), C.byref(cost), _ptr(r), _ptr(J), _ptr(psi), _ptr(rho1), _ptr(g)),
                 "dcs_evaluate")
        return dict(cost=cost.value, residuals=r, jacobians=J, psi=psi, rho1=rho1, gradient=g)

    def linearize(self, pose_xyt, want_gradient=True, out=None):
        x = np.ascontiguousarray(pose_xyt, dtype=np.float64)
        cost = C.c_double()
        g = out if out is not None else (np.empty((self.graph.n_poses, 3)) if want_gradient else None)
        self._ck(self.lib.dcs_linearize(self.h, _ptr(x), C.byref(cost), _ptr(g)), "dcs_linearize")
        return cost.value, g

    def linearize_resident(self, repeats=1, with_solver_setup=False):
        ms = C.c_float()
        self._ck(self.lib.dcs_linearize_resident(self.h, repeats, 1 if with_solver_setup else 0, C.byref(ms)), "dcs_linearize_resident")
        return ms.value

    def cost(self, pose_xyt=None):
        x = None if pose_xyt is None else np.ascontiguousarray(pose_xyt, dtype=np.float64)
        c = C.c_double()
        self._ck(self.lib.dcs_cost(self.h, _ptr(x), C.byref(c)), "dcs_cost")
        return c.value

    def pattern(self):
        nrows, nnzb = C.c_int32(), C.c_int32()
        self._ck(self.lib.dcs_get_pattern(self.h, C.byref(nrows), C.byref(nnzb), None, None), "dcs_get_pattern")
        rp = np.empty(nrows.value + 1, np.int32); ci = np.empty(nnzb.value, np.int32)
        self._ck(self.lib.dcs_get_pattern(self.h, C.byref(nrows), C.byref(nnzb), _ptr(rp), _ptr(ci)), "dcs_get_pattern")
        return rp, ci

    def hessian(self):
        rp, ci = self.pattern()
        v = np.empty((ci.shape[0], 3, 3))
        self._ck(self.lib.dcs_get_hessian(self.h, _ptr(v)), "dcs_get_hessian")
        return rp, ci, v

    def pcg_solve(self, lam, rhs):
        N = self.graph.n_poses
        lam = None if lam is None else np.ascontiguousarray(lam, dtype=np.float64).reshape(N, 3)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64).reshape(N, 3)
        w = np.empty((N, 3)); it = C.c_int32(); rel = C.c_double()
        self._ck(self.lib.dcs_pcg_solve(self.h, _ptr(lam), _ptr(rhs), _ptr(w), C.byref(it), C.byref(rel)), "dcs_pcg_solve")
        return w, it.value, rel.value

    def switches(self):
        """METHOD 2: switch value per edge after the last solve (1.0 on odometry edges)."""
        sw = np.empty(self.graph.n_edges)
        self._ck(self.lib.dcs_get_switches(self.h, _ptr(sw)), "dcs_get_switches")
        return sw

    def solve(self, pose_xyt=None):
        x = np.array(self.graph.pose_xyt if pose_xyt is None else pose_xyt, dtype=np.float64, order="C")
        s = Summary()
        cap = self.options.max_num_iterations + 2
        trace = (Iteration * cap)()
        self._ck(self.lib.dcs_solve(self.h, _ptr(x), C.byref(s), trace, cap), "dcs_solve")
        return x, s, [trace[i] for i in range(min(cap, s.num_iterations))]


def solve_batch(graphs, dcs_on=False, n_threads=8, return_poses=False, **opts):
    """dcs_solve_batch over a list of Graphs (N3: many tiny independent solves).  Returns (summaries, poses or None)."""
    lib = load_library()
    o = Options()
    lib.dcs_options_default(C.byref(o))
    o.dcs_on = 1 if dcs_on else 0
    for k, v in opts.items():
        setattr(o, k, v)
    n = len(graphs)
    items = (BatchItem * max(n, 1))()
    poses = [np.array(g.pose_xyt, dtype=np.float64, order="C") for g in graphs] if return_poses else None
    for i, g in enumerate(graphs):
        items[i].graph = _Graph(g.n_poses, g.n_edges, _ptr(g.pose_xyt), _ptr(g.edge_a), _ptr(g.edge_b), _ptr(g.meas_xyt),
                                _ptr(g.kind), g.fixed_pose)
        items[i].pose_xyt_inout = poses[i].ctypes.data if return_poses else None
    rc = lib.dcs_solve_batch(items, n, C.byref(o), n_threads)
    if rc:
        raise DcsError(rc, "dcs_solve_batch", lib.dcs_last_error().decode())
    sums = []
    for i in range(n):
        s = Summary()
        C.memmove(C.byref(s), C.byref(items[i].summary), C.sizeof(Summary))
        sums.append(s)
    return sums, poses

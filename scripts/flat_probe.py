import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
N = 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
s = D.Solver(g, dcs_on=True)
lib = D.load_library(); lib.dcs_debug_flat.restype = C.c_double; lib.dcs_debug_flat.argtypes = [C.c_void_p, C.c_int, C.c_int]
print("k_linearize us", 1e3 * s.linearize_resident(20) / 20)
for mode, name in ((0, "stream 32-B records only"), (1, "+ gather pose"), (3, "+ gather + owner 3x3 block stores (tile-interleaved compact)"),
                   (2, "records + owner stores, no gather")):
    print(f"mode {mode}: {name}: {lib.dcs_debug_flat(s.h, mode, 20):.1f} us")

lib.dcs_debug_atomic.restype = C.c_double; lib.dcs_debug_atomic.argtypes = [C.c_void_p, C.c_int, C.c_int]
for nper in (1, 3, 9):
    print(f"edge-centric scatter-add probe: {2*nper} fp64 atomics per edge x 4M edges: {lib.dcs_debug_atomic(s.h, nper, 10):.1f} us")

import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
N = 1_000_000
g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N))
s = D.Solver(g, dcs_on=True)
lib = D.load_library(); lib.dcs_debug_flat.restype = C.c_double; lib.dcs_debug_flat.argtypes = [C.c_void_p, C.c_int, C.c_int]
print("k_linearize us", 1e3 * s.linearize_resident(20) / 20)
for mode, name in ((0, "stream loads only (28 B x 8M)"), (1, "+ gather pose"), (3, "+ gather + owner-only 9 stores"), (7, "+ gather + all-lane 9 stores"), (2, "loads + owner stores, no gather"), (6, "loads + all stores, no gather")):
    print(f"mode {mode}: {name}: {lib.dcs_debug_flat(s.h, mode, 20):.1f} us")

for mode, name in ((11, 'gather + owner stores, tile-interleaved [slot/32][9][32]'), (15, 'gather + all-lane stores, tile-interleaved')):
    print(f'mode {mode}: {name}: {lib.dcs_debug_flat(s.h, mode, 20):.1f} us')
for kb in (75, 110):
    print(f'mode 11 tile-interleaved with {kb} KB smem: {lib.dcs_debug_flat(s.h, 11 | (kb << 8), 20):.1f} us')
# occupancy experiment: cap resident CTAs (256 threads = 8 warps each) with dynamic shared memory
for kb, ctas in ((0, 8), (56, 4), (75, 3), (110, 2), (200, 1)):
    print(f"mode 3 with {kb} KB smem/CTA (~{ctas} CTAs = {8*ctas} warps per SM): {lib.dcs_debug_flat(s.h, 3 | (kb << 8), 20):.1f} us")

lib.dcs_debug_atomic.restype = C.c_double; lib.dcs_debug_atomic.argtypes = [C.c_void_p, C.c_int, C.c_int]
for nper in (1, 3, 9):
    print(f"edge-centric scatter-add probe: {2*nper} fp64 atomics per edge x 4M edges: {lib.dcs_debug_atomic(s.h, nper, 10):.1f} us")

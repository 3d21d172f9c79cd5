"""Multi-rank PCG iteration time (weak scaling, 1M poses / 4M edges per GPU). Launch with torch.distributed.run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import torch, torch.distributed as dist
import dcs_b200 as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = torch.frombuffer(bytearray(D.nccl_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(buf, 0)
uid = bytes(buf.cpu().numpy().tobytes())
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
N = n * world
g = D.Graph.synthetic(N, int(2.7 * n) * world + 1, n_bogus=int(0.3 * n) * world)
s = D.Solver(g, dcs_on=True, device=local, rank=rank, world=world, nccl_unique_id=uid, max_num_iterations=1, pcg_max_iter=320,
             pcg_check_every=32, pcg_rel_tol=1e-30)
s.linearize_resident(3)
us = 1e3 * s.linearize_resident(10) / 10
x, sm, tr = s.solve()
if rank == 0:
    print(f"world {world}: N={N} E={g.n_edges} linearize {us:.1f} us/step, pcg {1e6 * sm.linear_solver_time_s / max(1, sm.total_pcg_iterations):.1f} us/iter "
          f"({sm.total_pcg_iterations} iterations)", flush=True)
import ctypes as C
lib = D.load_library(); out = (C.c_double * 4)()
lib.dcs_debug_comm.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.dcs_debug_comm(s.h, 50, out)
if rank == 0:
    print(f"world {world}: halo exchange {out[0]:.1f} us ({out[3]:.1f} MB sent per rank), allreduce(1) {out[1]:.1f} us, allreduce(2) {out[2]:.1f} us", flush=True)
out6 = (C.c_double * 6)()
lib.dcs_debug_pcg_stages.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.dcs_debug_pcg_stages(s.h, 50, out6)
if rank == 0:
    print(f"world {world}: stages us: spmv+fold {out6[0]:.1f}, allreduce(p.q) {out6[1]:.1f}, chain+fold {out6[2]:.1f}, allreduce(r.z) {out6[3]:.1f}, "
          f"direction {out6[4]:.1f}, halo {out6[5]:.1f}  sum {sum(out6):.1f}", flush=True)
s.close()
dist.barrier()
dist.destroy_process_group()

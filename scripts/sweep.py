"""BASELINE.json configs[4]: synthetic scaling sweep, 100 K - 20 M poses at 1 / 2 / 4 / 8 B200 (pose-range partition, halos
over NVLink).  Every size is the TOTAL graph, split over the ranks (strong scaling per size; the weak-scaling reading is
the diagonal).  Per size: eval+assembly step time, PCG iteration time (chain preconditioner, 320 iterations), create time.

  python scripts/sweep.py 1e5 1e6 1e7                                  # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 ... scripts/sweep.py 1e5 1e6 1e7 2e7

Writes gpurun_out/sweep_r02_w<world>.json (rank 0)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import dcs_b200 as D
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def uid():
    if not dist:
        return {}
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = torch.frombuffer(bytearray(D.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(buf, 0)
    return dict(nccl_unique_id=bytes(buf.cpu().numpy().tobytes()), rank=rank, world=world)


def rmax(x):
    if not dist:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


sizes = [int(float(x)) for x in (sys.argv[1:] or ["1e5", "3e5", "1e6", "3e6", "1e7"])]
out = []
for N in sizes:
    t = time.time(); g = D.Graph.synthetic(N, int(2.7 * N) + 1, n_bogus=int(0.3 * N)); tg = time.time() - t
    t = time.time()
    s = D.Solver(g, dcs_on=True, device=local, max_num_iterations=1, pcg_max_iter=320, pcg_check_every=32, pcg_rel_tol=1e-30, **uid())
    tc = time.time() - t
    s.linearize_resident(5)
    if dist: dist.barrier()
    us = rmax(1e3 * s.linearize_resident(20) / 20)
    x, sm, tr = s.solve()
    pcg = rmax(1e6 * sm.linear_solver_time_s / max(1, sm.total_pcg_iterations))
    rec = dict(n_gpus=world, n_poses=N, n_edges=g.n_edges, gen_s=round(tg, 2), create_s=round(rmax(tc), 3), linearize_us=round(us, 1),
               edges_per_s=g.n_edges / us * 1e6, frac_of_hbm_per_gpu=(108 * g.n_edges + 120 * N) / world / us / 1e3 / 6444.4,
               pcg_us_per_iter=round(pcg, 1), cost=tr[0].cost)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    out.append(rec)
    s.close(); del g
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_r02_w{world}.json"), "w"), indent=1)
if dist:
    dist.barrier()
    dist.destroy_process_group()

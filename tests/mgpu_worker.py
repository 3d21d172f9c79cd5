"""Worker for the multi-rank tests (launched with torch.distributed.run, one process per GPU).
Compares a world-size-N handle against a single-rank handle on the same graph."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "toy-robust-backend-slam_b200"))
import torch
import torch.distributed as dist

import dcs_b200 as D


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = torch.frombuffer(bytearray(D.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    g = D.Graph.synthetic(n, int(2.7 * n), n_bogus=int(0.3 * n))
    iters = 6
    s = D.Solver(g, dcs_on=True, device=local, rank=rank, world=world, nccl_unique_id=uid, max_num_iterations=iters)
    cost, grad = s.linearize(g.pose_xyt)
    row_lo, nrows = D.partition(g.n_poses, g.n_edges, rank, world)[:2]
    gl = torch.zeros(g.n_poses, 3, dtype=torch.float64, device="cuda")
    gl[row_lo:row_lo + nrows] = torch.from_numpy(grad[row_lo:row_lo + nrows]).cuda()
    dist.all_reduce(gl)
    c2 = s.cost(g.pose_xyt)
    x, summ, trace = s.solve()
    ok = True
    if rank == 0:
        with D.Solver(g, dcs_on=True, device=local, max_num_iterations=iters) as ref:
            c1, g1 = ref.linearize(g.pose_xyt)
            x1, s1, t1 = ref.solve()
        gm = np.abs(g1).max()
        checks = {
            "cost": abs(cost - c1) <= 1e-13 * c1,
            "cost_only": abs(c2 - c1) <= 1e-13 * c1,
            "gradient": np.abs(gl.cpu().numpy() - g1).max() <= 1e-12 * gm,
            "accept_sequence": [t.step_is_successful for t in trace] == [t.step_is_successful for t in t1],
            "final_cost": abs(summ.final_cost - s1.final_cost) <= 1e-9 * s1.final_cost,
            "poses": np.abs(x - x1).max() < 1e-6,
        }
        ok = all(checks.values())
        print("MGPU", world, "ranks:", checks, "final", summ.final_cost, s1.final_cost, "pcg", summ.total_pcg_iterations,
              s1.total_pcg_iterations, flush=True)
    s.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

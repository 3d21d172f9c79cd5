"""N>1 path.  CPU (gloo, world_size 2): the host-side pieces every rank runs — partition arithmetic and the
128-byte group id plumbing.  GPU (needs >= 2 devices): a 2-rank handle against a single-rank one."""
import os
import subprocess
import sys

import numpy as np
import pytest

import dcs_b200 as D
from conftest import ROOT


def test_partition_covers_rows_and_edges_exactly():
    for n, e, w in ((1228, 1533, 2), (1_000_000, 4_000_000, 8), (3500, 5553, 4), (5, 4, 3), (1024, 10, 1)):
        parts = [D.partition(n, e, r, w) for r in range(w)]
        assert sum(p[1] for p in parts) == n
        assert all(p[2] == parts[0][2] and p[2] % 1024 == 0 for p in parts)          # equal, window-aligned ranges
        assert [p[0] for p in parts] == [r * parts[0][2] for r in range(w)]
        assert parts[0][3] == 0 and parts[-1][4] == e
        assert all(parts[r][4] == parts[r + 1][3] for r in range(w - 1))
    with pytest.raises(D.DcsError):
        D.partition(10, 10, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what bench.py / the workers do before dcs_create: rank 0 owns the 128-byte id, everyone gets the same bytes
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.arange(128, dtype=torch.uint8) * 2 + 1
    dist.broadcast(buf, 0)
    n, e = 100_000, 400_000
    p = D.partition(n, e, rank, world)
    t = torch.tensor([p[1], p[4] - p[3]], dtype=torch.int64)
    dist.all_reduce(t)
    # timing plumbing of the bench: max over ranks
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    q.put((rank, bytes(buf.numpy().tobytes()), t.tolist(), ms.item()))
    dist.destroy_process_group()


def test_two_ranks_gloo_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0][1] == res[1][1] == bytes((np.arange(128) * 2 + 1).astype(np.uint8))
    assert res[0][2] == res[1][2] == [100_000, 400_000]
    assert res[0][3] == res[1][3] == 2.0


@pytest.mark.gpu
def test_two_gpu_ranks_match_single_rank():
    if D.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_worker.py"), "20000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MGPU 2 ranks" in p.stdout
    # once more on the -DDCS_CHECK build: device-side bounds asserts on the halo-indexed gathers
    lib = os.path.join(ROOT, "toy-robust-backend-slam_b200", "libdcs_b200_check.so")
    cmd[cmd.index("29517")] = "29521"
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, DCS_B200_LIB=lib))
    assert p.returncode == 0 and "DCS_CHECK failed" not in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.gpu
def test_four_gpu_ranks_match_single_rank():
    if D.device_count() < 4:
        pytest.skip("needs 4 GPUs (gpurun --gpus 4)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tests", "mgpu_worker.py"), "30000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MGPU 4 ranks" in p.stdout

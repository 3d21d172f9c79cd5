"""CPU suite: the CUDA translation unit cross-compiles for sm_100a and the hot kernels keep their resource
budget (no local-memory spills, register counts that give the occupancy the profiles were taken at).
A spill or a register jump in k_linearize / k_spmv is a silent 10-30 % regression, so it is a test."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "toy-robust-backend-slam_b200", "csrc")


@pytest.fixture(scope="module")
def ptxas_report():
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    out = subprocess.run(["make", "-C", CSRC, "ptxas-info"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    text = out.stdout + out.stderr
    rep = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads"
                         r"\n.*?Used (\d+) registers", text):
        rep[m.group(1)] = dict(stack=int(m.group(2)), spill_st=int(m.group(3)), spill_ld=int(m.group(4)), regs=int(m.group(5)))
    assert rep, text[-2000:]
    return rep


def _find(rep, needle):
    hits = {k: v for k, v in rep.items() if needle in k}
    assert hits, f"no kernel matching {needle}"
    return hits


# kernel (mangled-name fragment) -> register budget: k_linearize runs 16 one-warp CTAs per SM (128 registers),
# k_spmv / k_cost_rows 32 (64 registers); the chain kernels run less than one wave and may use the whole file.
BUDGET = {"11k_linearize": 128, "6k_spmv": 64, "11k_cost_rows": 80, "15k_pcg_direction": 40, "8k_expand": 40,
          "13k_pcg_cluster": 128}     # one 512-thread CTA per SM: 128 registers is the whole file


@pytest.mark.parametrize("needle", sorted(BUDGET))
def test_hot_kernels_do_not_spill_and_keep_their_registers(ptxas_report, needle):
    for name, r in _find(ptxas_report, needle).items():
        assert r["spill_st"] == 0 and r["spill_ld"] == 0, (name, r)
        assert r["regs"] <= BUDGET[needle], (name, r)


def test_no_kernel_spills(ptxas_report):
    # the per-edge parity dump (explicit 3x6 Jacobians) and the dev probes are not on the hot path
    allowed = ("k_edge_eval", "k_dbg_")
    spilled = {k: v for k, v in ptxas_report.items() if (v["spill_st"] or v["spill_ld"]) and not any(a in k for a in allowed)}
    assert not spilled, spilled

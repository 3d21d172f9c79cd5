#!/bin/bash
# round-2 GPU call G: parity suite after the tolerance fixes + LM trajectory sensitivity at 1 M poses
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > $O/r2_g_tests.log 2>&1
timeout 600 python scripts/lm_sens.py > $O/r2_g_sens.log 2>&1
tail -12 $O/r2_g_tests.log; cat $O/r2_g_sens.log

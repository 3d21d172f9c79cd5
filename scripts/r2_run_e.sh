#!/bin/bash
# round-2 GPU call E: the parity suite with METHOD 2, the batched solves, the bounds-assert build and the SYN10K full solve
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60) > $O/r2_e_tests.log 2>&1
tail -40 $O/r2_e_tests.log

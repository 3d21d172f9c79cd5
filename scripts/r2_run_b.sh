#!/bin/bash
# round-2 GPU call B: parity suite on the SELL layout, k_linearize variants, traffic floor, ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/r2_b_tests.log 2>&1
timeout 300 python scripts/variant_check.py > $O/r2_b_variants.log 2>&1
DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so timeout 200 python scripts/flat_probe.py > $O/r2_b_flat.log 2>&1
timeout 300 python scripts/prof_kernels.py > $O/r2_b_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_linearize|k_spmv|k_pcg_chain|k_cost_rows|k_expand" -c 12 -o $O/r2_b_prof -f python scripts/prof_kernels.py > $O/r2_b_ncu.log 2>&1
tail -5 $O/r2_b_tests.log; cat $O/r2_b_variants.log $O/r2_b_flat.log | tail -20; tail -3 $O/r2_b_prof_plain.log

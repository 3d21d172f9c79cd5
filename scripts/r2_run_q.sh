#!/bin/bash
# round-2 GPU call Q (1 GPU): chain kernel with the max shared-memory carve-out: PCG iteration time at 1 M poses, occupancy, tiny solves
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python scripts/prof_kernels.py > $O/r2_q_prof_plain.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_pcg_chain|k_spmv|k_pcg_direction|k_fold" -c 12 --csv --log-file $O/r2_q_chain.csv python scripts/prof_kernels.py > /dev/null 2>&1
timeout 200 python scripts/tiny_profile.py > $O/r2_q_tiny.log 2>&1
(timeout 600 python -m pytest tests -m gpu -q -k "bounds or method2 or pcg or chain or multi" 2>&1 | tail -5) > $O/r2_q_tests.log 2>&1
tail -3 $O/r2_q_prof_plain.log; head -1 $O/r2_q_tiny.log; tail -3 $O/r2_q_tests.log; grep -E "k_pcg_chain|k_spmv" $O/r2_q_chain.csv | cut -d, -f5,13- | head -30

#!/bin/bash
# round-2 GPU call K (1 GPU): final build (8-value blocks): parity suite, N=1 bench line, ncu launch list + --set full, 1-GPU sweep
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r2_k_tests.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_k_bench.json 2> $O/r2_k_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02.csv python bench.py --steps 5 --warmup 3 --lm-iters 1 --no-cpu > $O/r2_k_ncu_launch.log 2>&1
timeout 300 python scripts/prof_kernels.py > $O/r2_k_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_linearize|k_spmv|k_pcg_chain|k_cost_rows|k_expand|k_chain_factor" -c 14 -o $O/prof_r02_final -f python scripts/prof_kernels.py > $O/r2_k_ncu.log 2>&1
(timeout 400 python scripts/sweep.py 1e5 3e5 1e6 3e6 1e7 2e7 2>&1 | grep -E "^\{|Error|error") > $O/r2_k_sweep1.log 2>&1
tail -6 $O/r2_k_tests.log; tail -c 1500 $O/r2_k_bench.json; tail -3 $O/r2_k_bench.err; tail -3 $O/r2_k_prof_plain.log; cat $O/r2_k_sweep1.log

#!/bin/bash
# round-2 GPU call D: parity suite, N=1 bench line (full LM solve), ncu launch list + --set full on the shipped build
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/r2_d_tests.log 2>&1
timeout 1200 python bench.py --steps 20 --warmup 5 > $O/r2_d_bench.json 2> $O/r2_d_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02.csv python bench.py --steps 5 --warmup 3 --lm-iters 1 --no-cpu > $O/r2_d_ncu_launch.log 2>&1
timeout 300 python scripts/prof_kernels.py > $O/r2_d_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_linearize|k_spmv|k_pcg_chain|k_cost_rows|k_expand|k_chain_factor" -c 14 -o $O/prof_r02_final -f python scripts/prof_kernels.py > $O/r2_d_ncu.log 2>&1
tail -5 $O/r2_d_tests.log; tail -c 2500 $O/r2_d_bench.json; tail -5 $O/r2_d_bench.err; tail -3 $O/r2_d_prof_plain.log

#!/bin/bash
# round-2 GPU call X (1 GPU): first contact of the one-launch cluster PCG (k_pcg_cluster): both paths side by side on the
# reference's datasets, then the small-graph part of the parity suite (which now runs through the cluster kernel)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 150 python scripts/small_pcg.py > $O/r2_x_small.log 2> $O/r2_x_small.err
echo "small_pcg rc=$?"
(timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pcg_matches or full_lm_solve or reproducible or method2 or batched or bounds or chain_precond or single_edge or hub or duplicate or drop_in" 2>&1 | tail -15) > $O/r2_x_tests.log 2>&1
cat $O/r2_x_small.log; tail -5 $O/r2_x_small.err; tail -8 $O/r2_x_tests.log

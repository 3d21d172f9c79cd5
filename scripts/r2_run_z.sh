#!/bin/bash
# round-2 GPU call Z (1 GPU): k_pcg_cluster v2 (fp64 paired factors, three predicated rounds in flight, r / w in shared
# memory): both PCG paths side by side, phase cycles from the dev library, small-graph part of the parity suite
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T=${1:-z}
timeout -s KILL 150 python scripts/small_pcg.py > $O/r2_${T}_small.log 2> $O/r2_${T}_small.err; echo "small_pcg rc=$?"
python - <<PY
import json
for l in open('$O/r2_${T}_small.log'):
    d=json.loads(l); ls=d.get('linear_solve',{}); lm=d['lm']
    print(d['case'], d['options'], 'ok', d['ok'], 'iters', ls.get('iters_cluster'), ls.get('iters_general'), 'us/it', round(ls.get('us_per_iter_cluster',0),2), round(ls.get('us_per_iter_general',0),2),
          'lm pcg', lm['pcg_iterations_cluster'], lm['pcg_iterations_general'], 'us/pcg', round(lm['us_per_pcg_cluster'],2), round(lm['us_per_pcg_general'],2), 'cost_rel', lm['max_cost_rel_diff'], 'final', lm['final_cost_cluster'])
PY
tail -3 $O/r2_${T}_small.err
DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so timeout -s KILL 100 python scripts/small_pcg_prof.py > $O/r2_${T}_phases.log 2>&1
cat $O/r2_${T}_phases.log
(timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pcg_matches or full_lm_solve or reproducible or method2 or batched or bounds or chain_precond or single_edge or hub or duplicate or drop_in" 2>&1 | tail -15) > $O/r2_${T}_tests.log 2>&1
tail -4 $O/r2_${T}_tests.log

#!/bin/bash
# round-2 GPU call (1 GPU), short: both PCG paths side by side + phase cycles of k_pcg_cluster (no pytest)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T=${1:-s}
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

#!/bin/bash
# round-2 GPU call U (1 GPU): FINAL build: parity suite, N=1 bench line, ncu launch list + --set full (the profiles/r02_* evidence)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r2_u_tests.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_u_bench.json 2> $O/r2_u_bench.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02.csv python bench.py --steps 5 --warmup 3 --lm-iters 1 --no-cpu --no-extras > $O/r2_u_ncu_launch.log 2>&1
timeout 300 python scripts/prof_kernels.py > $O/r2_u_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_linearize|k_spmv|k_pcg_chain|k_cost_rows|k_expand|k_chain_factor" -c 14 -o $O/prof_r02_final -f python scripts/prof_kernels.py > $O/r2_u_ncu.log 2>&1
tail -4 $O/r2_u_tests.log; tail -3 $O/r2_u_bench.err; tail -2 $O/r2_u_prof_plain.log; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2_u_bench.json').read().strip().splitlines()[-1])
print('value',b['value'],'frac',b['roofline']['frac'],'lm',b['lm']['seconds'],b['lm']['us_per_pcg_iteration'], b['lm']['final_cost'], b['lm']['max_true_residual'])
e=b['extras']; print(e['method2']['seconds'], e['method2']['us_per_pcg_iteration']); print(e['batched_tiny_solves'])
PY

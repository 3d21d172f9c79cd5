#!/bin/bash
# round-2 GPU call P (1 GPU): the 512-thread chain-preconditioner kernel: parity suite, PCG iteration time, tiny-solve profile, bench line
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > $O/r2_p_tests.log 2>&1
timeout 200 python scripts/prof_kernels.py > $O/r2_p_prof_plain.log 2>&1
timeout 200 python scripts/tiny_profile.py > $O/r2_p_tiny.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_p_bench.json 2> $O/r2_p_bench.err
echo "bench rc=$?"
tail -6 $O/r2_p_tests.log; tail -3 $O/r2_p_prof_plain.log; head -2 $O/r2_p_tiny.log; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2_p_bench.json').read().strip().splitlines()[-1])
print('value',b['value'],'frac',b['roofline']['frac'],'lm',b['lm']['seconds'],b['lm']['us_per_pcg_iteration'], b['lm']['final_cost'], b['lm']['max_true_residual'])
print(json.dumps(b['extras'])[:1500])
PY

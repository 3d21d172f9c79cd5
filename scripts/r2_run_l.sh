#!/bin/bash
# round-2 GPU call L (1 GPU): parity suite + the bench line with the METHOD 2 / batched-solve side measurements, create-time breakdown
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r2_l_tests.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_l_bench.json 2> $O/r2_l_bench.err
echo "bench rc=$?"
DCS_CREATE_TIMING=1 timeout 200 python scripts/sweep.py 1e6 > $O/r2_l_create.log 2>&1
tail -6 $O/r2_l_tests.log; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2_l_bench.json').read().strip().splitlines()[-1])
print('value',b['value'],'frac',b['roofline']['frac'],'lm',b['lm']['seconds'],b['lm']['us_per_pcg_iteration'])
print(json.dumps(b['extras'])[:3000])
PY
tail -3 $O/r2_l_bench.err; grep dcs_create $O/r2_l_create.log

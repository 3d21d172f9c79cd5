#!/bin/bash
# round-2 final GPU call (1 GPU): smoke(), the full GPU parity suite and the default bench line on the final build
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/r2_f_smoke.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > $O/r2_f_tests.log 2>&1
timeout 400 python bench.py > $O/r2_f_bench.json 2> $O/r2_f_bench.err
echo "bench rc=$?"
cat $O/r2_f_smoke.log; tail -4 $O/r2_f_tests.log; tail -3 $O/r2_f_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2_f_bench.json').read().strip().splitlines()[-1])
print('value',b['value'],'ms',b['ms_per_step'],'frac',b['roofline']['frac'],'e2e',b['e2e']['value'],'lm s',b['lm']['seconds'],b['lm']['us_per_pcg_iteration'],b['lm']['final_cost'])
x=b['extras']; print('m2',x['method2']['seconds'],x['method2']['us_per_pcg_iteration']); print('batch',{k:v for k,v in x['batched_tiny_solves'].items() if k.startswith('threads')})
print('small',json.dumps(x['small_graph_pcg']))
PY

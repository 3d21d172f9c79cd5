#!/bin/bash
# round-2 GPU call W (2 GPUs): last build: smoke(), full parity suite (incl. 2-rank on both builds), N=2 bench line
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/r2_w_smoke.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > $O/r2_w_tests.log 2>&1
timeout 600 $TR --master-port 29661 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_w_bench2.json 2> $O/r2_w_bench2.err
echo "bench rc=$?"
cat $O/r2_w_smoke.log; tail -3 $O/r2_w_tests.log; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2_w_bench2.json').read().strip().splitlines()[-1])
print('value',b['value'],'ms',b['ms_per_step'],'lm',b['lm']['seconds'],b['lm']['us_per_pcg_iteration'], 'parity', b['parity']['ok'], 'strong', b['strong_1m']['ms_per_step'], b['strong_1m']['us_per_pcg_iteration'], b['strong_1m']['seconds'], b['strong_1m']['parity']['ok'], b['strong_1m']['parity']['leading_iterations_agreeing_1e-9'], b['config']['create_s'])
PY

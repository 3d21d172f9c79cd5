#!/bin/bash
# round-2 GPU call M (2 GPUs): multi-rank parity on the final build, create-time breakdown, size sweep at 2 GPUs, the N=2 bench line
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q 2>&1 | tail -12) > $O/r2_m_tests.log 2>&1
(DCS_CREATE_TIMING=1 timeout 300 $TR --master-port 29631 scripts/sweep.py 1e5 1e6 1e7 2e7 2>&1 | grep -E "^\{|dcs_create|Error|error") > $O/r2_m_sweep2.log 2>&1
timeout 900 $TR --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_m_bench2.json 2> $O/r2_m_bench2.err
echo "bench rc=$?"
tail -4 $O/r2_m_tests.log; cat $O/r2_m_sweep2.log | cut -c1-400; tail -c 1800 $O/r2_m_bench2.json; tail -3 $O/r2_m_bench2.err

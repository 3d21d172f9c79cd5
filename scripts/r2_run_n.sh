#!/bin/bash
# round-2 GPU call N (4 GPUs): 2- and 4-rank parity (both builds) after the create-time change, size sweep at 4 GPUs with the create-time breakdown
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q 2>&1 | tail -12) > $O/r2_n_tests.log 2>&1
(DCS_CREATE_TIMING=1 timeout 300 $TR --master-port 29641 scripts/sweep.py 1e6 1e7 2e7 2>&1 | grep -E "^\{|dcs_create|Error|error") > $O/r2_n_sweep4.log 2>&1
tail -4 $O/r2_n_tests.log; grep -E "^\{|sell layout|stream/nccl|pose upload" $O/r2_n_sweep4.log | cut -c1-330

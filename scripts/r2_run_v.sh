#!/bin/bash
# round-2 GPU call V (2 GPUs): final build: 2-rank parity on both builds, PCG exchange-step times (weak 1 M per GPU, strong 1 M total)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q 2>&1 | tail -6) > $O/r2_v_tests.log 2>&1
export DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so
(timeout 200 $TR --master-port 29651 scripts/mgpu_pcg.py 2>&1 | grep -E "world|Error|error") > $O/r2_v_pcg_weak.log 2>&1
(timeout 200 $TR --master-port 29652 scripts/mgpu_pcg.py 5e5 2>&1 | grep -E "world|Error|error") > $O/r2_v_pcg_strong.log 2>&1
tail -3 $O/r2_v_tests.log; cat $O/r2_v_pcg_weak.log $O/r2_v_pcg_strong.log

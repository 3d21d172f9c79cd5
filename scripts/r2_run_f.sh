#!/bin/bash
# round-2 GPU call F (2 GPUs): multi-rank parity (+ bounds-assert build), PCG exchange-step times with the peer-memory
# scalar exchange vs ncclAllReduce, push variants, and the N=2 bench line (weak + strong leg + parity block)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q 2>&1 | tail -30) > $O/r2_f_tests.log 2>&1
export DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so
(timeout 300 $TR --master-port 29601 scripts/mgpu_pcg.py 2>&1 | grep -E "world|Error|error") > $O/r2_f_pcg_xchg.log 2>&1
(DCS_SCALARS=nccl timeout 300 $TR --master-port 29602 scripts/mgpu_pcg.py 2>&1 | grep -E "world|Error|error") > $O/r2_f_pcg_nccl.log 2>&1
(DCS_PUSH_PER_THREAD=1 timeout 300 $TR --master-port 29603 scripts/mgpu_pcg.py 2>&1 | grep -E "world|Error|error") > $O/r2_f_pcg_push1.log 2>&1
unset DCS_B200_LIB
timeout 1500 $TR --master-port 29604 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_f_bench2.json 2> $O/r2_f_bench2.err
echo "bench rc=$?"
tail -8 $O/r2_f_tests.log; cat $O/r2_f_pcg_*.log; tail -c 3000 $O/r2_f_bench2.json; tail -5 $O/r2_f_bench2.err

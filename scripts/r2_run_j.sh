#!/bin/bash
# round-2 GPU call J (8 GPUs): size sweep, PCG exchange-step times, the N=8 bench line (weak + strong leg + parity)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 240 $TR --master-port 29621 scripts/sweep.py 1e6 1e7 2e7 2>&1 | grep -E "^\{|Error|error" ) > $O/r2_j_sweep8.log 2>&1
export DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so
(timeout 120 $TR --master-port 29622 scripts/mgpu_pcg.py 2>&1 | grep -E "world|Error|error") > $O/r2_j_pcg8.log 2>&1
unset DCS_B200_LIB
timeout 400 $TR --master-port 29623 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_j_bench8.json 2> $O/r2_j_bench8.err
echo "bench rc=$?"
cat $O/r2_j_sweep8.log $O/r2_j_pcg8.log; tail -c 2500 $O/r2_j_bench8.json; tail -5 $O/r2_j_bench8.err

#!/bin/bash
# round-2 GPU call I (2 GPUs): full parity suite (incl. 2-rank parity on both builds) with the halo-last row order, the
# two-pass SpMV overlapped with the halo push, the peer-memory scalar exchange and the 256-bit push stores; PCG stage times
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r2_i_tests.log 2>&1
export DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so
(timeout 300 $TR --master-port 29611 scripts/mgpu_pcg.py 2>&1 | grep -E "world|Error|error") > $O/r2_i_pcg_a_overlap.log 2>&1
(DCS_OVERLAP=0 timeout 300 $TR --master-port 29612 scripts/mgpu_pcg.py 2>&1 | grep -E "world|Error|error") > $O/r2_i_pcg_b_nooverlap.log 2>&1
(timeout 300 $TR --master-port 29613 scripts/mgpu_pcg.py 5e5 2>&1 | grep -E "world|Error|error") > $O/r2_i_pcg_c_strong.log 2>&1
tail -8 $O/r2_i_tests.log; cat $O/r2_i_pcg_*.log

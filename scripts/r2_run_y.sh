#!/bin/bash
# round-2 GPU call Y (1 GPU): phase breakdown of k_pcg_cluster (dev library timers) + one ncu --set full capture with source
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout -s KILL 150 python scripts/small_pcg.py > $O/r2_y_small.log 2> $O/r2_y_small.err; echo "small_pcg rc=$?"; grep -o '"ok": [a-z]*' $O/r2_y_small.log | tr '\n' ' '; echo
DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so timeout -s KILL 100 python scripts/small_pcg_prof.py > $O/r2_y_phases.log 2>&1
cat $O/r2_y_phases.log
timeout -s KILL 100 python scripts/small_pcg_prof.py M3500_100_seed1 > $O/r2_y_plain.log 2>&1 && \
timeout -s KILL 240 ncu --set full --import-source on --clock-control none -k regex:k_pcg_cluster -c 2 -f -o $O/r2_y_cluster python scripts/small_pcg_prof.py M3500_100_seed1 > $O/r2_y_ncu.log 2>&1
tail -3 $O/r2_y_ncu.log; ls -la $O/r2_y_cluster.ncu-rep

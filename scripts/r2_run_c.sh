#!/bin/bash
# round-2 GPU call C: k_linearize prefetch variants, traffic floor, the new bench line (N=1)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 python scripts/variant_check.py > $O/r2_c_variants.log 2>&1
DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libdcs_b200_dev.so timeout 200 python scripts/flat_probe.py > $O/r2_c_flat.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_c_bench.json 2> $O/r2_c_bench.err
cat $O/r2_c_variants.log $O/r2_c_flat.log | tail -24; tail -c 3000 $O/r2_c_bench.json; tail -5 $O/r2_c_bench.err

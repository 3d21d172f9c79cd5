#!/bin/bash
# round-2 GPU call (1 GPU): product-loop variants of k_pcg_cluster (dev builds with phase timers): phase cycles + wall time.
# Historical: the libvar_dev_p*u*.so libraries were built from a compile-time switch (product loop form x rounds in flight, on
# top of the fp64-factor variant) that was removed with the variant; results: profiles/r02_small_graphs.md section 3.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in p1u2 p1u3 p1u1 p2u3; do
  echo "== $v"
  DCS_B200_LIB=$PWD/toy-robust-backend-slam_b200/libvar_dev_$v.so timeout -s KILL 60 python scripts/small_pcg_prof.py 2>&1 | tee -a $O/r2_variants.log
done

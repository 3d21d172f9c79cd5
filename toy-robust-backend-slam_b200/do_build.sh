#!/bin/sh
# do_build.sh DATASET NUM_OUTLIER_LOOPS METHOD   — same command line and the same save/ build/ data/ drawer/ flow as
# the reference's DCS-ceres/do_build.sh:1-16 (clean save/, build, run ./main from build/ so that ../data and
# ../save resolve, then plot).  Differences: save/ is created first (the reference's `cd save; rm -rf *` wipes the
# source tree on a fresh clone), the build is two Makefiles (nvcc sm_100a + g++) instead of cmake+Ceres+Eigen+boost,
# and the plot step is skipped when drawer/ (the reference's matplotlib scripts, unchanged consumers of save/*.txt)
# is not present.   Data: put DATASET.g2o under data/ (or export DCS_DATA_PATH).
set -e
here=$(cd "$(dirname "$0")" && pwd)
cd "$here"
mkdir -p save build data
rm -rf save/*
make -s -C csrc
make -s -C host
cp host/main build/main
cd build
LD_LIBRARY_PATH="$here:$LD_LIBRARY_PATH" ./main "$1" "$2" "$3"
cd ..
if [ -x drawer/do_plot.sh ]; then
  cd drawer && ./do_plot.sh && cd ..
fi

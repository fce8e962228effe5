#!/bin/bash
# ncu captures of the tree rollout kernel (rolled instantiation) at 65536 and 1024 particles
set -u
OUT=gpurun_out/${1:-r02o}
mkdir -p $OUT
for K in 65536 1024; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_tree -c 1 -s 3 -o $OUT/tree_$K python tools/bench_tree.py --sizes $K > /dev/null 2>> $OUT/log.txt; echo "ncu $K exit $?" | tee -a $OUT/log.txt
done

#!/bin/bash
# ncu capture of the tree kernel on the half-cheetah (contacts in every substep)
set -u
OUT=gpurun_out/${1:-r02v}
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_tree -c 1 -s 3 -o $OUT/tree_cheetah_8192 python tools/bench_tree.py --model cheetah --sizes 8192 > /dev/null 2>> $OUT/log.txt; echo "ncu exit $?" | tee -a $OUT/log.txt

#!/bin/bash
# first hardware run of the tree rollout kernel (K11): parity tests, sizes, one ncu capture
set -u
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tree_gpu.py -m gpu -q > $OUT/tests.log 2>&1; echo "tree tests exit $?" | tee -a $OUT/log.txt
timeout 600 python tools/bench_tree.py --sizes 1024,8192,65536 > $OUT/bench_tree.jsonl 2>> $OUT/log.txt; echo "bench exit $?" | tee -a $OUT/log.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_tree -c 1 -s 3 -o $OUT/tree_65536 python tools/bench_tree.py --sizes 65536 > /dev/null 2>> $OUT/log.txt; echo "ncu exit $?" | tee -a $OUT/log.txt
timeout 300 python examples/run_mpc.py --config examples/configs/swimmer-v0.yml --controller mppi --n_episodes 1 --cuda_graph > $OUT/run_mpc_swimmer.log 2>&1; echo "example exit $?" | tee -a $OUT/log.txt
tail -3 $OUT/tests.log; cat $OUT/bench_tree.jsonl; tail -4 $OUT/run_mpc_swimmer.log

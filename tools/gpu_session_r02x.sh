#!/bin/bash
# final build: ncu launch list of the bench command (per-kernel share of the step) + one full capture of K1 (DRAM traffic)
set -u
OUT=gpurun_out/${1:-r02x}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > $OUT/ncu_bench.log 2>&1; echo "launch list exit $?" | tee -a $OUT/log.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_reacher_kernel -s 3 -c 1 -o $OUT/k1_full_65536 \
    python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_k1.log 2>&1; echo "k1 capture exit $?" | tee -a $OUT/log.txt
timeout 300 python bench.py --steps 200 --warmup 10 > $OUT/bench.json 2>> $OUT/log.txt

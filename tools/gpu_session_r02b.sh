#!/bin/bash
# Round 2, second GPU call: the role-split rollout kernel -- parity on hardware, timing against the
# one-thread-per-particle kernel at the per-GPU particle counts of N = 2..32 ranks, one full ncu capture.
set -u
OUT=gpurun_out/r02b
mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/log.txt; timeout 900 "$@" >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt; }
run python -m pytest tests/test_rollout_split_gpu.py -m gpu -x -q
run python -m pytest tests -m gpu -x -q
for K in 2048 8192 16384 32768; do
  for S in 0 1048576; do
    MJB_SPLIT_MAX_K=$S timeout 600 python tools/k1_variants.py one $K 2>> $OUT/log.txt | sed "s/^{/{\"split_max_k\": $S, /" >> $OUT/k1_split.jsonl
  done
done
MJB_SPLIT_MAX_K=1048576 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:rollout_reacher_split_kernel -s 6 -c 1 -o $OUT/k1_split_full_8192 python tools/k1_variants.py one 8192 > $OUT/ncu_k1_split_8192.log 2>&1
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_rollout_split_gpu.py -m gpu -q -x -k "matches_oracle and 77" > $OUT/racecheck_split.log 2>&1; echo "racecheck split exit $?" >> $OUT/log.txt
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_rollout_split_gpu.py -m gpu -q -x -k "matches_oracle and 77 or single_particle or per_worker" > $OUT/memcheck_split.log 2>&1; echo "memcheck split exit $?" >> $OUT/log.txt
tail -12 $OUT/log.txt; cut -c1-330 $OUT/k1_split.jsonl

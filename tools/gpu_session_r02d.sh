#!/bin/bash
# split kernel v2 (angle-addition sin/cos, no 4th barrier, 64-lane blocks beyond 148 groups): parity + timing + timeline
set -u
OUT=gpurun_out/r02d
mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/log.txt; timeout 900 "$@" >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt; }
run python -m pytest tests/test_rollout_split_gpu.py tests/test_rollout_gpu.py -m gpu -x -q
for K in 1024 2048 4096 8192 16384; do
  for S in 0 1048576; do
    MJB_SPLIT_MAX_K=$S timeout 600 python tools/k1_variants.py one $K 2>> $OUT/log.txt | sed "s/^{/{\"split_max_k\": $S, /" >> $OUT/k1_split.jsonl
  done
done
for K in 2048 8192; do python tools/split_timeline.py run $K >> $OUT/timeline.jsonl 2>> $OUT/log.txt; done
MJB_SPLIT_MAX_K=1048576 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:rollout_reacher_split_kernel -s 6 -c 1 -o $OUT/k1_split_full_8192 python tools/k1_variants.py one 8192 > $OUT/ncu_k1_split_8192.log 2>&1
tail -4 $OUT/log.txt; python - <<'P'
import json
for l in open("gpurun_out/r02d/k1_split.jsonl"):
    r = json.loads(l); print(r["K"], "split" if r["split_max_k"] else "mono ", min(r["ms_min"]), "%.1e" % r["rel_err_vs_oracle"])
P
cat $OUT/timeline.jsonl

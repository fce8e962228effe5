#!/bin/bash
# role-split kernel with the small-code changes (rolled recursions in roles 0/1, shared factor code for roles 1/2)
set -u
OUT=gpurun_out/r02j
mkdir -p $OUT
timeout 600 python -m pytest tests/test_rollout_split_gpu.py tests/test_efc_pin.py -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests exit $?"
for K in 1024 2048 4096 8192; do
  MJB_SPLIT_MAX_K=1048576 timeout 600 python tools/k1_variants.py one $K 2>> $OUT/log.txt | sed "s/^{/{\"build\": \"rolled+shared\", /" >> $OUT/k1_split.jsonl
  MJB_LIB_PATH=gpurun_variants/lib_split_unrolled.so MJB_SPLIT_MAX_K=1048576 timeout 600 python tools/k1_variants.py one $K 2>> $OUT/log.txt | sed "s/^{/{\"build\": \"unrolled+shared\", /" >> $OUT/k1_split.jsonl
done
for K in 2048 8192; do python tools/split_timeline.py run $K >> $OUT/timeline.jsonl 2>> $OUT/log.txt; done
MJB_SPLIT_MAX_K=1048576 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:rollout_reacher_split_kernel -s 6 -c 1 -o $OUT/k1_split_full_8192 python tools/k1_variants.py one 8192 > $OUT/ncu.log 2>&1
tail -2 $OUT/tests.log; python - <<'P'
import json
for l in open("gpurun_out/r02j/k1_split.jsonl"):
    r = json.loads(l); print(r["K"], r["build"], min(r["ms_min"]), "%.1e" % r["rel_err_vs_oracle"])
P
cat $OUT/timeline.jsonl

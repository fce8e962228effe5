#!/bin/bash
# rolled vs unrolled mass/bias in the thread-per-particle kernel: parity, timing at 8192 / 16384 / 65536, footprint
set -u
OUT=gpurun_out/r02i
mkdir -p $OUT
for K in 8192 16384 65536; do
  for V in default advance; do
    MJB_SPLIT_MAX_K=0 MJB_LIB_PATH=gpurun_variants/lib_$V.so timeout 600 python tools/k1_variants.py one $K 2>> $OUT/log.txt >> $OUT/k1_advance.jsonl
  done
done
MJB_SPLIT_MAX_K=0 MJB_LIB_PATH=gpurun_variants/lib_default.so timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:rollout_reacher_kernel -s 6 -c 1 -o $OUT/k1_advance_65536 python tools/k1_variants.py one 65536 > $OUT/ncu.log 2>&1
python - <<'P'
import json
for l in open("gpurun_out/r02i/k1_advance.jsonl"):
    r = json.loads(l); print(r["K"], r["variant"], min(r["ms_min"]), "%.1e" % r["rel_err_vs_oracle"])
P
tail -3 $OUT/log.txt

#!/bin/bash
# three instantiations of K1 (throughput / latency-oriented / role-split) over the per-GPU particle counts of N = 2..64 ranks
set -u
OUT=gpurun_out/r02e
mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/log.txt; timeout 900 "$@" >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt; }
run python -m pytest tests -m gpu -x -q
for K in 1024 4096 8192 16384 32768 65536; do
  for M in "0 0 throughput" "0 1048576 latency" "1048576 0 split"; do
    set -- $M
    [ "$3" = "split" ] && [ $K -gt 16384 ] && continue
    MJB_SPLIT_MAX_K=$1 MJB_LAT_MAX_K=$2 timeout 600 python tools/k1_variants.py one $K 2>> $OUT/log.txt | sed "s/^{/{\"kernel\": \"$3\", /" >> $OUT/k1_kernels.jsonl
  done
done
tail -4 $OUT/log.txt; python - <<'P'
import json
for l in open("gpurun_out/r02e/k1_kernels.jsonl"):
    r = json.loads(l); print(r["K"], r["kernel"], min(r["ms_min"]), "%.1e" % r["rel_err_vs_oracle"])
P

#!/bin/bash
# fused update tail: GPU tests, bench (graph + eager), launch list
set -u
OUT=gpurun_out/r02f
mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/log.txt; timeout 900 "$@" >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt; }
run python -m pytest tests -m gpu -q
timeout 600 python bench.py --steps 300 --warmup 10 > $OUT/bench.json 2>> $OUT/log.txt
timeout 600 python bench.py --steps 300 --warmup 10 --no-graph --no-cpu-baseline > $OUT/bench_nograph.json 2>> $OUT/log.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > $OUT/ncu_bench.log 2>&1
grep -E "passed|failed" $OUT/log.txt | tail -3; python - <<'P'
import json
for f in ("bench.json", "bench_nograph.json"):
    b = json.load(open("gpurun_out/r02f/" + f)); print(f, b["ms_per_step"], b["e2e"]["ms_per_step"], b["breakdown"])
P
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02f/launches.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
seq = [(r[ki][:60], float(r[vi].replace(",", ""))) for r in rows[hi + 1:] if len(r) == len(hdr)]
for n, v in seq[-12:]: print("%-62s %10.1f ns" % (n, v))
P

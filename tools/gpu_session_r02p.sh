#!/bin/bash
# tree rollout kernel: planar instantiation vs the general one (swimmer), parity tests, one ncu capture of the planar kernel
set -u
OUT=gpurun_out/${1:-r02p}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tree_gpu.py -m gpu -q > $OUT/tests.log 2>&1; echo "tree tests exit $?" | tee -a $OUT/log.txt
timeout 300 python tools/bench_tree.py --sizes 1024,8192,65536 >> $OUT/bench_tree.jsonl 2>> $OUT/log.txt
timeout 300 python tools/bench_tree.py --sizes 1024,8192,65536 --no-planar >> $OUT/bench_tree.jsonl 2>> $OUT/log.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_tree -c 1 -s 3 -o $OUT/tree_planar_65536 python tools/bench_tree.py --sizes 65536 > /dev/null 2>> $OUT/log.txt; echo "ncu exit $?" | tee -a $OUT/log.txt
tail -3 $OUT/tests.log
python - <<P
import json
for l in open("$OUT/bench_tree.jsonl"):
    d = json.loads(l); print(d["instantiation"], d["num_particles"], "kernel %.3f ms  step %.3f  e2e %.3f err %.1e" % (d["rollout_kernel_ms"], d["mpc_step_ms"], d["e2e_ms"], d["rel_err_vs_oracle"]))
P

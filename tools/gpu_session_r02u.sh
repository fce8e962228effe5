#!/bin/bash
set -u
OUT=gpurun_out/${1:-r02u}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tree_gpu.py -m gpu -q > $OUT/tests.log 2>&1; echo "tree tests exit $?" | tee -a $OUT/log.txt
for M in swimmer swimmer-nocontact cheetah; do
  timeout 300 python tools/bench_tree.py --model $M --sizes 1024,8192,65536 >> $OUT/bench_tree.jsonl 2>> $OUT/log.txt
done
tail -3 $OUT/tests.log
python - <<P
import json
for l in open("$OUT/bench_tree.jsonl"):
    d = json.loads(l); print(d["model"], d["instantiation"], d["num_particles"], "kernel %.3f ms  step %.3f  e2e %.3f err %.1e" % (d["rollout_kernel_ms"], d["mpc_step_ms"], d["e2e_ms"], d["rel_err_vs_oracle"]))
P

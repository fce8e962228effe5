#!/bin/bash
# tree rollout kernel: per-link loops unrolled (variant 1) vs rolled with compile-time nv (0) vs fully run-time (2)
set -u
OUT=gpurun_out/${1:-r02m}
mkdir -p $OUT
for v in 0 1 2; do
  MJB_TREE_VARIANT=$v timeout 300 python tools/bench_tree.py --sizes 1024,8192,65536 2>> $OUT/log.txt | sed "s/^{/{\"variant\": $v, /" >> $OUT/tree_variants.jsonl
done
timeout 300 python -m pytest tests/test_tree_gpu.py -m gpu -q 2>&1 | tail -2
python - <<P
import json
for l in open("$OUT/tree_variants.jsonl"):
    d = json.loads(l); print(d["variant"], d["num_particles"], "kernel %.3f ms  step %.3f  err %.1e" % (d["rollout_kernel_ms"], d["mpc_step_ms"], d["rel_err_vs_oracle"]))
P

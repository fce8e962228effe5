#!/bin/bash
# overlapped noise kernel: forked before the rollout (shares the GPU with K1) vs after it (shares it with the update kernels)
set -u
OUT=gpurun_out/${1:-r02z}
mkdir -p $OUT
for rep in 1 2; do
for mode in early late; do
  MJB_NOISE_FORK=$mode timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>> $OUT/log.txt | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', b['ms_per_step'], b['e2e']['ms_per_step'], b['breakdown'])" | tee -a $OUT/fork.txt
done
done

#!/bin/bash
# the strong-scaling series of the bench on one 8-GPU box, final build:  gpurun --gpus 8 -- 'bash tools/gpu_session_final_scale.sh'
set -u
OUT=gpurun_out/r02_final_scale
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --steps 500 --warmup 10 > $OUT/bench_n1.json 2>> $OUT/log.txt
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 29553 bench.py --gpus $n --steps 500 --warmup 10 > $OUT/bench_n$n.json 2>> $OUT/log.txt
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_final_scale/bench_*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.4f e2e %.4f" % (b["ms_per_step"], b["e2e"].get("ms_per_step", 0)), "k1", b.get("roofline", {}).get("ms_per_launch"), "frac", b.get("roofline", {}).get("frac"), b.get("sharded_parity", {}).get("max_rel") if b.get("sharded_parity") else None)
    except Exception as e:
        print(f, "unreadable", e)
P

#!/bin/bash
# 8-GPU box: the non-headline configurations at full size (sharded), the 1024-instance sweep partitioned over the GPUs,
# and the strong-scaling series of the bench.   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_session_final8.sh'
set -u
OUT=gpurun_out/r02_final8
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python tools/bench_configs.py > $OUT/configs_n1.jsonl 2>> $OUT/log.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29551 tools/bench_configs.py > $OUT/configs_n8.jsonl 2>> $OUT/log.txt
for C in mppi pfmpc; do
  timeout 300 python examples/run_sweep.py --instances 1024 --steps 100 --controller $C > $OUT/sweep_${C}_n1.json 2>> $OUT/log.txt
  timeout 300 $TR --nproc-per-node 8 --master-port 29552 examples/run_sweep.py --instances 1024 --steps 100 --controller $C > $OUT/sweep_${C}_n8.json 2>> $OUT/log.txt
done
timeout 600 python bench.py --steps 500 --warmup 10 > $OUT/bench_n1.json 2>> $OUT/log.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/log.txt
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 29553 bench.py --gpus $n --steps 500 --warmup 10 > $OUT/bench_n$n.json 2>> $OUT/log.txt
done
grep -v "^\*\|OMP_NUM" $OUT/log.txt | tail -5
for f in $OUT/configs_n1.jsonl $OUT/configs_n8.jsonl; do echo $f; python - $f <<'P'
import json, sys
for l in open(sys.argv[1]):
    try:
        r = json.loads(l); print("  %-72s %8.4f ms  %s" % (r.get("config", "?")[:72], r.get("ms_per_step", float("nan")), r.get("error", "")))
    except Exception:
        pass
P
done
for f in $OUT/sweep_*.json; do echo $f; tail -1 $f | cut -c1-400; done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_final8/bench_*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.4f e2e %.4f" % (b["ms_per_step"], b["e2e"].get("ms_per_step", 0) if isinstance(b.get("e2e"), dict) else 0), "k1", b.get("roofline", {}).get("ms_per_launch"), b.get("sharded_parity"))
    except Exception as e:
        print(f, "unreadable", e)
P

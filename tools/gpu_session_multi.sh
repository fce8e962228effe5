#!/bin/bash
# multi-GPU verification on N GPUs of one box (gpurun --gpus N): sharded-vs-unsharded checks of all five controllers on
# real NCCL ranks + the NVLink peer-memory exchange, then the bench at every power of two up to N.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_session_multi.sh N'
set -u
N=${1:-2}
OUT=gpurun_out/r02_multi_n$N
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.csv 2>&1
echo "== pytest tests/test_multigpu_gpu.py" | tee -a $OUT/log.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_zz_native_step_gpu.py -m gpu -q >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt
for n in 2 4 8; do
  [ $n -gt $N ] && continue
  for K in 2048 65536; do
    echo "== multigpu_check N=$n K=$K" | tee -a $OUT/multigpu_check.log
    MJB_CHECK_K=$K timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
        tests/helpers/multigpu_check.py >> $OUT/multigpu_check.log 2>&1; echo "   exit $?" | tee -a $OUT/multigpu_check.log
  done
done
timeout 600 python bench.py --steps 300 --warmup 10 > $OUT/bench_n1.json 2>> $OUT/log.txt
for n in 2 4 8; do
  [ $n -gt $N ] && continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 \
      bench.py --gpus $n --steps 300 --warmup 10 > $OUT/bench_n$n.json 2>> $OUT/log.txt
done
grep -E "ok \(N|FAIL|exit" $OUT/multigpu_check.log | tail -40
python - <<P
import json, glob
for f in sorted(glob.glob("$OUT/bench_n*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f e2e %.4f k1 %.4f" % (b["ms_per_step"], b["e2e"]["ms_per_step"], b["roofline"]["ms_per_launch"]), b["breakdown"], b.get("sharded_parity"), b.get("exchange"))
    except Exception as e:
        print(f, "unreadable", e)
P
tail -5 $OUT/log.txt

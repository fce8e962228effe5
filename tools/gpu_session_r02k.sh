#!/bin/bash
# after the host-I/O graphs + fast single-state path: multi-GPU tests on 2 real ranks, bench at N = 1, 2
set -u
OUT=gpurun_out/r02k
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_zz_native_step_gpu.py tests/test_api_gpu.py tests/test_episode_gpu.py -m gpu -q > $OUT/tests.log 2>&1; echo "tests exit $?" | tee -a $OUT/log.txt
MJB_CHECK_K=65536 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/helpers/multigpu_check.py > $OUT/multigpu_check.log 2>&1; echo "check exit $?" | tee -a $OUT/log.txt
timeout 600 python bench.py --steps 500 --warmup 10 > $OUT/bench_n1.json 2>> $OUT/log.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 500 --warmup 10 > $OUT/bench_n2.json 2>> $OUT/log.txt
tail -3 $OUT/tests.log; grep -cE "ok \(N" $OUT/multigpu_check.log; grep -E "FAIL" $OUT/multigpu_check.log | head
python - <<P
import json, glob
for f in sorted(glob.glob("$OUT/bench_n*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f e2e %.4f k1 %.4f" % (b["ms_per_step"], b["e2e"]["ms_per_step"], b["roofline"]["ms_per_launch"]), b.get("sharded_parity"))
    except Exception as e:
        print(f, "unreadable", e)
P

#!/bin/bash
# tree family with contacts: GPU tests, memcheck of the contact paths, sizes for swimmer (with / without self-contact) and cheetah
set -u
OUT=gpurun_out/${1:-r02t}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tree_gpu.py -m gpu -q > $OUT/tests.log 2>&1; echo "tree tests exit $?" | tee -a $OUT/log.txt
for M in swimmer swimmer-nocontact cheetah; do
  timeout 300 python tools/bench_tree.py --model $M --sizes 1024,8192,65536 >> $OUT/bench_tree.jsonl 2>> $OUT/log.txt
done
timeout 300 python tools/bench_tree.py --model swimmer-nocontact --sizes 65536 --no-planar >> $OUT/bench_tree.jsonl 2>> $OUT/log.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tree_gpu.py -m gpu -q -x -k "contact or walker or cheetah_rollouts" > $OUT/memcheck_contacts.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/log.txt
timeout 300 python examples/run_mpc.py --config examples/configs/half_cheetah-v0.yml --controller mppi --n_episodes 1 --cuda_graph > $OUT/run_mpc_cheetah.log 2>&1; echo "cheetah example exit $?" | tee -a $OUT/log.txt
timeout 300 python examples/run_mpc.py --config examples/configs/swimmer-v0.yml --controller mppi --n_episodes 1 --cuda_graph > $OUT/run_mpc_swimmer.log 2>&1; echo "swimmer example exit $?" | tee -a $OUT/log.txt
tail -3 $OUT/tests.log; tail -3 $OUT/memcheck_contacts.log; tail -3 $OUT/run_mpc_cheetah.log; tail -3 $OUT/run_mpc_swimmer.log
python - <<P
import json
for l in open("$OUT/bench_tree.jsonl"):
    d = json.loads(l); print(d["model"], d["instantiation"], d["num_particles"], "kernel %.3f ms  step %.3f  e2e %.3f err %.1e" % (d["rollout_kernel_ms"], d["mpc_step_ms"], d["e2e_ms"], d["rel_err_vs_oracle"]))
P

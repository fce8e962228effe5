#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke(), bench (both arms), plus sanitizer passes on the kernels of round 2
set -u
OUT=gpurun_out/r02_validate
mkdir -p $OUT
echo "== pytest -m gpu" | tee -a $OUT/log.txt; timeout 900 python -m pytest tests -m gpu -x -q >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt
echo "== smoke" | tee -a $OUT/log.txt; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt
timeout 600 python bench.py > $OUT/bench.json 2>> $OUT/log.txt; echo "bench exit $?" | tee -a $OUT/log.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/log.txt; echo "reference exit $?" | tee -a $OUT/log.txt
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_controllers_gpu.py -m gpu -q -x -k "elite_select_cluster or resample_certified or batched_rs" > $OUT/memcheck_update.log 2>&1; echo "memcheck update kernels exit $?" | tee -a $OUT/log.txt
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_controllers_gpu.py tests/test_noise_gpu.py -m gpu -q -x -k "elite_select_cluster and 4096 or resample_certified and 4096 or noise" > $OUT/racecheck_update.log 2>&1; echo "racecheck update/noise kernels exit $?" | tee -a $OUT/log.txt
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_native_step_gpu.py -m gpu -q -x -k "bit_identical and mppi" > $OUT/memcheck_step.log 2>&1; echo "memcheck native step exit $?" | tee -a $OUT/log.txt
grep -E "passed|failed|exit|smoke ok" $OUT/log.txt | tail -12; head -c 700 $OUT/bench.json; echo; head -c 400 $OUT/bench_reference.json

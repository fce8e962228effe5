#!/bin/bash
# One GPU-box call that produces everything the next round needs first (tests, bench lines, A/B timings of the
# changes made without a GPU, launch list, one full ncu capture of the rollout kernel).  Build the variant
# libraries HERE first:   python tools/k1_variants.py build
# then:   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh'          (about 15 GPU-minutes)
# Everything lands in gpurun_out/session/; copy what should be judged into profiles/.
set -u
OUT=gpurun_out/session
mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/log.txt; "$@" >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt; }

run python -m pytest tests -m gpu -x -q
python bench.py --steps 300 --warmup 10 > $OUT/bench.json 2>> $OUT/log.txt
python bench.py --steps 300 --warmup 10 --impl reference --cpu-particles 8192 > $OUT/bench_reference.json 2>> $OUT/log.txt
# eager step through the one-call native step vs the step-by-step Python path (host time between launches)
python bench.py --steps 300 --warmup 10 --no-graph --no-cpu-baseline > $OUT/bench_nograph_native.json 2>> $OUT/log.txt
MJB_FUSED_STEP=0 python bench.py --steps 300 --warmup 10 --no-graph --no-cpu-baseline > $OUT/bench_nograph_stepwise.json 2>> $OUT/log.txt
# ... and with the next step's noise drawn on a side stream during the rollout (results bit-identical; DESIGN 4.1 queue item 2)
python bench.py --steps 300 --warmup 10 --no-graph --overlap-noise --no-cpu-baseline > $OUT/bench_nograph_native_overlap.json 2>> $OUT/log.txt
# rollout kernel: current build vs without the rank-one repair vs the previous commit; also the under-filled sizes
if [ -d gpurun_variants ]; then
  python tools/k1_variants.py run 65536 > $OUT/k1_variants_65536.jsonl 2>> $OUT/log.txt
  python tools/k1_variants.py run 8192 > $OUT/k1_variants_8192.jsonl 2>> $OUT/log.txt
fi
python tools/bench_configs.py > $OUT/configs.jsonl 2>> $OUT/log.txt
python examples/run_sweep.py --instances 1024 --steps 100 > $OUT/sweep_1gpu.json 2>> $OUT/log.txt
# per-kernel launch list of the bench command (cold-cache, serialised: shares, not absolutes) and one full capture
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_reacher_kernel -s 3 -c 1 -o $OUT/k1_full \
    python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:noise_kernel -s 3 -c 1 -o $OUT/k2_full \
    python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_k2.log 2>&1
# memory and shared-memory race checks on small problems (the emulator cannot see either)
compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > $OUT/memcheck_smoke.log 2>&1; echo "memcheck smoke exit $?" >> $OUT/log.txt
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_controllers_gpu.py -m gpu -q -x -k "mppi_update or cem_update or dmd_update or pfmpc or random_shooting" > $OUT/racecheck_updates.log 2>&1; echo "racecheck updates exit $?" >> $OUT/log.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.csv 2>&1
tail -5 $OUT/log.txt; head -c 600 $OUT/bench.json

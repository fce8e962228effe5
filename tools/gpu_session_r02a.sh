#!/bin/bash
# Round 2, first GPU call: where does the time go at 8192 particles per GPU (lone-warp regime), is the rank-one
# repair paying, how slow are CEM / RS / PFMPC at K=65536, is any MuJoCo on the box.
# Build the variants HERE first: python tools/k1_variants.py build
set -u
OUT=gpurun_out/r02a
mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/log.txt; timeout 600 "$@" >> $OUT/log.txt 2>&1; echo "   exit $?" | tee -a $OUT/log.txt; }
{ python -c "import mujoco; print('mujoco', mujoco.__version__)"; python -c "import mujoco_py; print('mujoco_py ok')"; python -c "import dm_control; print('dm_control ok')";
  find / \( -iname '*mujoco*' -o -iname 'mjpro*' -o -iname 'mjkey*' \) -not -path '/proc/*' -not -path '*/gpurun*' 2>/dev/null | head -20; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'muj|dm_control|pinocchio|bullet' ; echo "probe done"; } > $OUT/mujoco_probe.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.csv 2>&1
run python -m pytest tests -m gpu -x -q
timeout 600 python bench.py --steps 300 --warmup 10 > $OUT/bench.json 2>> $OUT/log.txt
for K in 65536 8192 2048; do
  timeout 900 python tools/k1_variants.py run $K > $OUT/k1_variants_$K.jsonl 2>> $OUT/log.txt
done
timeout 900 python tools/bench_configs.py --only fullsize > $OUT/configs_fullsize.jsonl 2>> $OUT/log.txt
# lone-warp regime: one full capture of the production rollout kernel at 8192 particles
MJB_LIB_PATH=gpurun_variants/lib_default.so timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:rollout_reacher_kernel -s 6 -c 1 -o $OUT/k1_full_8192 python tools/k1_variants.py one 8192 > $OUT/ncu_k1_8192.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_reacher_kernel -s 3 -c 1 -o $OUT/k1_full_65536 \
    python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
tail -5 $OUT/log.txt; cat $OUT/k1_variants_8192.jsonl | cut -c1-400; head -c 400 $OUT/bench.json

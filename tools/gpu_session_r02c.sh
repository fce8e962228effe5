#!/bin/bash
# footprint probe: the thread-per-particle kernel with ~28 KB hot loop (fake_small: wrong physics) vs the real 39 KB one
set -u
OUT=gpurun_out/r02c
mkdir -p $OUT
for K in 8192 65536; do
  for V in default fake_small; do
    MJB_SPLIT_MAX_K=0 MJB_LIB_PATH=gpurun_variants/lib_$V.so timeout 600 python tools/k1_variants.py one $K 2>> $OUT/log.txt >> $OUT/k1_fake.jsonl
    MJB_SPLIT_MAX_K=0 MJB_LIB_PATH=gpurun_variants/lib_$V.so timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:rollout_reacher_kernel -s 6 -c 1 -o $OUT/k1_${V}_$K python tools/k1_variants.py one $K > $OUT/ncu_${V}_$K.log 2>&1
  done
done
cut -c1-300 $OUT/k1_fake.jsonl

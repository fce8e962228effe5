#!/bin/bash
# tree family after per-worker models: GPU tests + compute-sanitizer memcheck over the parity tests
set -u
OUT=gpurun_out/${1:-r02s}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tree_gpu.py -m gpu -q > $OUT/tests.log 2>&1; echo "tree tests exit $?" | tee -a $OUT/log.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tree_gpu.py -m gpu -q -x -k "oracle or worker or limits or planar" > $OUT/memcheck_tree.log 2>&1; echo "memcheck tree exit $?" | tee -a $OUT/log.txt
tail -3 $OUT/tests.log; tail -4 $OUT/memcheck_tree.log

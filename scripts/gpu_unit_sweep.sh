#!/bin/bash
# Runs every GPU kernel test in its own process (a trapped kernel poisons its CUDA context) and
# collects the logs under gpurun_out/.  Usage: gpurun -- bash scripts/gpu_unit_sweep.sh [pytest file]
FILE=${1:-tests/test_gpu_kernels.py}
mkdir -p gpurun_out
OUT=gpurun_out/unit_sweep.log
: > $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $OUT 2>&1
# one process per test function (all its parametrisations together)
TESTS=$(python -m pytest $FILE -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
for t in $TESTS; do
  echo "=== $t" >> $OUT
  timeout 300 python -m pytest "$t" -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -25 >> $OUT
done
grep -E "^(=== |[0-9]+ passed|[0-9]+ failed|FAILED|ERROR|.*passed|.*failed)" $OUT | tail -120

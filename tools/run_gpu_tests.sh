#!/bin/bash
# Runs every -m gpu test in its own process (a trapped kernel poisons only its own context) and collects a summary
# under gpurun_out/.  Usage: bash tools/run_gpu_tests.sh [pytest -k expression]
mkdir -p gpurun_out
OUT=gpurun_out/gpu_tests.log
: > $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $OUT 2>&1
ids=$(python -m pytest tests -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::")
pass=0; fail=0
for id in $ids; do
  echo "=== $id" >> $OUT
  timeout 300 python -m pytest "$id" -x -q -s -p no:cacheprovider >> $OUT 2>&1
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id"; else fail=$((fail+1)); echo "FAIL($rc) $id"; fi
done
echo "passed=$pass failed=$fail" | tee -a $OUT

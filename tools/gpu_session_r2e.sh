#!/bin/bash
mkdir -p gpurun_out
for sz in "2 128 128" "4 128 128" "2 256 256"; do
  timeout 300 python tools/sanitize_step.py tf32 $sz 2>&1 | tail -2
  echo "plain tf32 $sz rc=$?"
done
ONET_TRACE_JOINS=1 timeout 300 python tools/sanitize_step.py tf32 4 128 128 2>&1 | tail -12
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "tf32" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_model_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2e_model.log 2>&1
echo "model rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2e_model.log | tail -8
grep -E "teacher-forced:|tf32 B=|vs tf32-emul" gpurun_out/r2e_model.log | cut -c1-1000

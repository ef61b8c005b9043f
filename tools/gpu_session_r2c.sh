#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -s -p no:cacheprovider -k "tf32" > gpurun_out/r2c_kernels.log 2>&1
echo "kernels rc=$?"; grep -E "passed|failed|tf32 operands:" gpurun_out/r2c_kernels.log | tail -3; grep -E "^FAILED" gpurun_out/r2c_kernels.log | head
timeout 900 python -m pytest tests/test_parity_gpu.py -q -s -p no:cacheprovider -k "tf32" > gpurun_out/r2c_parity.log 2>&1
echo "parity rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c_parity.log | tail -5
grep -E "teacher-forced:|tf32 B=|vs tf32-emul" gpurun_out/r2c_parity.log | cut -c1-1100

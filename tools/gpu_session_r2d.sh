#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "tf32 or bn_relu or first_layer" > gpurun_out/r2d_kernels.log 2>&1
echo "kernels rc=$?"; grep -E "passed|failed" gpurun_out/r2d_kernels.log | tail -2; grep -E "^FAILED" gpurun_out/r2d_kernels.log | head -20
grep -E "^E  +Assert|^E  +assert" gpurun_out/r2d_kernels.log | sort | uniq -c | head
ONET_TRACE_CALLS=1 timeout 300 python tools/sanitize_step.py tf32 2 128 128 2>&1 | tail -4
ONET_TRACE_CALLS=1 timeout 300 python tools/sanitize_step.py tf32 2 32 32 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_step.py tf32 2 64 64 > gpurun_out/r2d_memcheck_tf32.log 2>&1
echo "memcheck tf32 rc=$?"; grep -E "Invalid|at .*kernel|ERROR SUMMARY|rep " gpurun_out/r2d_memcheck_tf32.log | head -12
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider -k "not tf32" > gpurun_out/r2d_model.log 2>&1
echo "model rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2d_model.log | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k, v in list(d['per_kernel'].items())[:14]:
    print(k, v)
PY

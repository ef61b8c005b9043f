#!/bin/bash
# round-2 session B: TF32 tensor-core kernels - kernel tests first (own process each: a trap poisons only its own context)
mkdir -p gpurun_out
for t in "tests/test_kernels_gpu.py::test_conv3x3_tf32_fwd_dgrad_wgrad" "tests/test_kernels_gpu.py::test_tf32_operand_semantics" "tests/test_kernels_gpu.py::test_convT2x2"; do
  timeout 300 python -m pytest "$t" -q -s -p no:cacheprovider -k "tf32" > gpurun_out/r2b_$(basename ${t##*::}).log 2>&1
  echo "$t rc=$?"; grep -E "passed|failed|tf32 operands" gpurun_out/r2b_$(basename ${t##*::}).log | tail -3
  grep -E "^FAILED|Error|error" gpurun_out/r2b_$(basename ${t##*::}).log | head -12
done
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_infer_gpu.py tests/test_model_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2b_parity.log 2>&1
echo "parity rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2b_parity.log | tail -12
grep -E "teacher-forced|tf32 B=" gpurun_out/r2b_parity.log | cut -c1-900
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'])
print('modes', {m: (r['value'], r['e2e']['value'], r['peak_memory_gb']) for m, r in d['modes'].items()})
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "first_layer" 2>&1 | tail -15 | cut -c1-300
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py tests/test_training_trajectory_gpu.py tests/test_full_size_properties_gpu.py -q -p no:cacheprovider > gpurun_out/r2k_tests.log 2>&1
echo "model rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2k_tests.log | tail -8; grep -E "^E  " gpurun_out/r2k_tests.log | head -12 | cut -c1-250
for v in "" "ONET_NO_FIRST_GRAM=1"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
  python - "$v" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print(f"[{sys.argv[1]:22s}] value {d['value']:.1f} ms {d['ms_per_step']:.3f} e2e {d['e2e']['value']:.1f} clocks {d['clocks']['sm_mhz']}", {k: v['ms_per_step'] for k, v in d['per_kernel'].items() if 'first' in k})
PY
done

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_parity_gpu.py -q -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2g_tests.log | tail -8
for v in "" "ONET_NO_HEAD_FUSION=1"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2g_bench_$v.json 2> gpurun_out/r2g_bench.err
  echo "bench [$v] rc=$?"; tail -1 gpurun_out/r2g_bench.err | cut -c1-200
  python - "$v" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2g_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('   value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'clocks', d['clocks']['sm_mhz'], d['clocks']['power_w_median'])
print('   ', {k: v['ms_per_step'] for k, v in d['per_kernel'].items() if 'head' in k or 'bn_relu' in k or 'first' in k})
PY
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "convT or head" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py tests/test_infer_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'])
print({k: (v['ms_per_step'], v['calls']) for k, v in d['per_kernel'].items() if 'tapgemm' in k or 'head' in k})
print({k: v for k, v in d['roofline']['shapes'].items()} if 'tapgemm' in d['roofline']['kernel'] else '')
PY
python tools/profile_layer.py convT 128 128 128 128 64 5
python tools/profile_layer.py convT 128 64 64 256 128 5

#!/bin/bash
# tests + bench after the window-kernel grid change, then the ncu evidence of the final code
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2u_tests.log; grep -n "^E  \|passed\|failed" gpurun_out/r2u_tests.log | cut -c1-300 | head
python bench.py --no-extra --no-profile --no-cpu-baseline --steps 30 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print("value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY
bash tools/gpu_session_r2_ncu.sh 2>&1 | tail -25

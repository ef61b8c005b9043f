#!/bin/bash
# first conv on the warp-level tensor-core path: tests, then same-box A/B of the whole step
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2l_tests.log
cat gpurun_out/r2l_tests.log
for v in mma nomma mma nomma; do
  if [ $v = nomma ]; then export ONET_NO_FIRST_MMA=1; else unset ONET_NO_FIRST_MMA; fi
  python bench.py --no-extra --no-profile --steps 30 > gpurun_out/r2l_bench_$v.json 2> gpurun_out/r2l_bench_$v.err
  python - $v <<'PY'
import json, sys
v=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2l_bench_{v}.json').read().strip().splitlines()[-1])
print(v, "value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY
done

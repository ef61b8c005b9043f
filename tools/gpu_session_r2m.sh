#!/bin/bash
# cost of the BatchNorm-statistics epilogue: forward (stats) vs dgrad form (no stats) of the same conv shapes; tests; bench
mkdir -p gpurun_out
out=gpurun_out/r2m_stats_cost.txt; rm -f $out
for s in "128 128 128 64 128" "128 128 128 128 128" "128 64 64 128 256" "128 256 256 64 64" "128 128 128 256 128" "128 256 256 128 64"; do
  python tools/profile_layer.py fwd $s 5 >> $out 2>&1
  python tools/profile_layer.py dgrad $s 5 >> $out 2>&1
done
cat $out
python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2m_tests.log; grep -n "^E  \|passed\|failed" gpurun_out/r2m_tests.log | cut -c1-300 | head
python bench.py --no-extra --no-profile --steps 30 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print("value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY

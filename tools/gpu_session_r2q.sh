#!/bin/bash
# A/B: BatchNorm-backward reduce folded into the dgrad epilogue for narrower layers, after the epilogue rework
mkdir -p gpurun_out
for c in 256 128 64 256 128; do
  ONET_BNRED_MIN_C=$c python bench.py --no-extra --no-profile --no-cpu-baseline --steps 30 > gpurun_out/r2q_bench_$c.json 2> gpurun_out/r2q_bench_$c.err
  python - $c <<'PY'
import json, sys
v=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2q_bench_{v}.json').read().strip().splitlines()[-1])
print("min_c", v, "value", round(d['value'],1), "ms", round(d['ms_per_step'],3), d['clocks']['sm_mhz'])
PY
done

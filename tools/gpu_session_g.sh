#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_a.json 2>/dev/null
ONET_NO_2CTA_WGRAD=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_b.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c.json 2>/dev/null
python - <<'PY'
import json
for n in ("ab_a","ab_b","ab_c"):
    d=json.load(open(f"gpurun_out/{n}.json")); print(n, round(d["value"],1), round(d["ms_per_step"],2), d["e2e"]["value"], d["clocks"], round(sum(v["ms_per_step"] for v in d["kernels"].values()),2))
d=json.load(open("gpurun_out/ab_c.json")); [print(k, v) for k,v in d['kernels'].items()]
PY

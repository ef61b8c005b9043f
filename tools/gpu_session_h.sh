#!/bin/bash
# A/B inside one box: CTA-pair conv kernels on/off, graph replay
mkdir -p gpurun_out
for i in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_2cta_$i.json 2>/dev/null
  ONET_NO_2CTA=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_1cta_$i.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("ab_2cta_1","ab_1cta_1","ab_2cta_2","ab_1cta_2"):
    d=json.load(open(f"gpurun_out/{n}.json")); print(n, round(d["value"],1), round(d["ms_per_step"],2), d["clocks"], round(sum(v["ms_per_step"] for v in d["kernels"].values()),2))
PY

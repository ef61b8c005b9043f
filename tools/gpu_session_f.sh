#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ONET_BENCH_DETAIL=gpurun_out/detail.tsv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_a.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c.json 2>/dev/null
python - <<'PY'
import json
for n in ("ab_a","ab_c"):
    d=json.load(open(f"gpurun_out/{n}.json")); print(n, round(d["value"],1), round(d["ms_per_step"],2), d["e2e"]["value"], d["clocks"]["sm_mhz"], round(sum(v["ms_per_step"] for v in d["kernels"].values()),2))
PY
sort -t$'\t' -k3 -n -r gpurun_out/detail.tsv | head -6

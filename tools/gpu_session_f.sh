#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "wgrad or convT" 2>&1 | tail -3
for cfg in "wgrad 128 64 64 256 256" "wgrad 128 32 32 512 512" "wgrad 128 32 32 1024 512" "wgrad 128 128 128 256 128" "wgrad 128 16 16 1024 1024" "wgrad 128 128 128 128 128"; do
  timeout 120 python tools/profile_layer.py $cfg 5 2>&1 | tail -1
  ONET_WG_KS_FASTEST=1 timeout 120 python tools/profile_layer.py $cfg 5 2>&1 | tail -1 | sed 's/^/   (old order) /'
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_a.json 2>/dev/null
ONET_WG_KS_FASTEST=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_b.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c.json 2>/dev/null
python - <<'PY'
import json
for n in ("ab_a","ab_b","ab_c"):
    d=json.load(open(f"gpurun_out/{n}.json")); print(n, round(d["value"],1), round(d["ms_per_step"],2), d["e2e"]["value"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_median"], round(sum(v["ms_per_step"] for v in d["kernels"].values()),2))
PY

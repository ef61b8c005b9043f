#!/bin/bash
# two-stream backward: correctness, then A/B against the serial schedule on the same box
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -5
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "bn_relu" 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$name.json")); print("$name", round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms  e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"))
except Exception as e: print("$name failed", e)
PY
}
run serial ONET_NO_WGRAD_OVERLAP=1
run overlap ONET_X=1
run overlap_b148 ONET_BN_BWD_BLOCKS=148
run overlap_b296 ONET_BN_BWD_BLOCKS=296
run serial2 ONET_NO_WGRAD_OVERLAP=1
run overlap2 ONET_X=1
tail -3 gpurun_out/ab_overlap.err

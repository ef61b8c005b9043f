#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "fused_bn_backward" 2>&1 | tail -3
for r in 1 2; do for cfg in off 64 128 256; do
unset ONET_NO_BNRED_FUSION ONET_BNRED_MIN_C
if [ $cfg = off ]; then export ONET_NO_BNRED_FUSION=1; else export ONET_BNRED_MIN_C=$cfg; fi
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fuse_$cfg.json 2> gpurun_out/bench_fuse_$cfg.err; echo "bench fuse=$cfg rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_fuse_$cfg.json")); pk=d["per_kernel"]
print("fuse=$cfg", round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms  e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], "bn_bwd", pk.get("bn_relu_bwd",{}).get("ms_per_step"), "apply", pk.get("bn_relu_bwd_apply",{}).get("ms_per_step"), "dgrad_bnred", pk.get("conv3x3_dgrad_bnred",{}).get("ms_per_step"), pk.get("conv3x3_dgrad_bnred",{}).get("calls"))
PY
done; done

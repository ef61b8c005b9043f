#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --workload zy3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zy3.json 2> gpurun_out/bench_zy3.err; echo "zy3 rc=$?"; tail -2 gpurun_out/bench_zy3.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_zy3.json")); print("zy3", round(d["value"],1), round(d["ms_per_step"],2), d["e2e"]["value"], d["config"]["workload"][:80]); [print(k,v) for k,v in list(d['kernels'].items())[:12]]
PY

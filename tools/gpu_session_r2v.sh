#!/bin/bash
# closed-form first conv for in_chns = 3 (bf16): kernel tests, all tests, ZY-3 shape bench with per-call times, headline bench
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "first_layer" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2v_tests_first.log; grep -n "^E  \|passed\|failed" gpurun_out/r2v_tests_first.log | cut -c1-300 | head
python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2v_tests.log; grep -n "^E  \|passed\|failed" gpurun_out/r2v_tests.log | cut -c1-300 | head
ONET_BENCH_DETAIL=gpurun_out/r2v_step_detail_zy3.tsv python bench.py --workload zy3 --no-extra --no-cpu-baseline --steps 10 > gpurun_out/r2v_bench_zy3.json 2> gpurun_out/r2v_bench_zy3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v_bench_zy3.json").read().strip().splitlines()[-1]); print("zy3", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
for l in open("gpurun_out/r2v_step_detail_zy3.tsv"):
    if "first" in l: print(l.strip()[:110])
PY
python bench.py --no-extra --no-profile --no-cpu-baseline --steps 30 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print("value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY

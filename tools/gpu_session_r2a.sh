#!/bin/bash
# round-2 session A: all GPU tests (with the new emulated-oracle parity tests), smoke, the bench line with its sub-records
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity.json
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2a_tests.log | cut -c1-400
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2a_tests.log | tail -20
python __graft_entry__.py smoke 2>&1 | tail -3
ONET_BENCH_DETAIL=gpurun_out/r2a_detail.tsv python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])
print('hbm', d['roofline_hbm'])
print('infer', {k: d['infer'][k] for k in ('value', 'e2e', 'roofline')}, d['infer'].get('cpu_baseline'))
print('zy3', d['zy3']['value'], d['zy3']['e2e'], d['zy3']['peak_memory_gb'])
print('modes', {m: (r['value'], r['e2e']['value']) for m, r in d['modes'].items()})
print('cpu', d.get('cpu_baseline'))
PY

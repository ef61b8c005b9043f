#!/bin/bash
# Regenerates the single-GPU evidence set of a round in one gpurun call: all GPU tests (parity values -> gpurun_out/r2_parity.json),
# smoke, both bench arms, the per-call dump and the per-layer roofline table.  Multi-GPU and ncu evidence: tools/gpu_session_r2_n2.sh,
# _n4.sh, _n8.sh, _ncu.sh.  Copy what should be judged from gpurun_out/ to profiles/.
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity.json
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1
echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2_tests.log | tail -6
python __graft_entry__.py smoke 2>&1 | tail -2
ONET_BENCH_DETAIL=gpurun_out/r2_step_detail.tsv python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err | cut -c1-200
python tools/roofline_table.py gpurun_out/r2_step_detail.tsv > gpurun_out/r2_per_layer_roofline.md; tail -5 gpurun_out/r2_per_layer_roofline.md
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'step frac', round(d['step_frac_of_sustained_peak'], 3))
print('roof', d['roofline']['kernel'], round(d['roofline']['achieved'], 1), round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'])
h = d['roofline_hbm']; print('hbm', h['kernel'], round(h['achieved']), round(h['frac'], 3), 'traffic', h['traffic'])
print('infer', round(d['infer']['value'], 1), round(d['infer']['e2e']['value'], 1), round(d['infer']['roofline']['frac'], 3), d['infer'].get('cpu_baseline', {}).get('value'))
print('zy3', round(d['zy3']['value'], 1), round(d['zy3']['e2e']['value'], 1))
print('modes', {m: (round(r['value'], 1), round(r['e2e']['value'], 1)) for m, r in d['modes'].items()})
print('cpu', d.get('cpu_baseline'))
PY

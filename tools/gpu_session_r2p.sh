#!/bin/bash
# head forward on warp-level MMAs: tests, whole-step A/B on one box, per-call times
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2p_tests.log; grep -n "^E  \|passed\|failed" gpurun_out/r2p_tests.log | cut -c1-300 | head
for v in mma nomma mma nomma; do
  if [ $v = nomma ]; then export ONET_NO_HEAD_MMA=1; else unset ONET_NO_HEAD_MMA; fi
  python bench.py --no-extra --no-profile --steps 30 > gpurun_out/r2p_bench_$v.json 2> gpurun_out/r2p_bench_$v.err
  python - $v <<'PY'
import json, sys
v=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2p_bench_{v}.json').read().strip().splitlines()[-1])
print(v, "value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY
done
unset ONET_NO_HEAD_MMA
ONET_BENCH_DETAIL=gpurun_out/r2p_step_detail.tsv python bench.py --no-extra --no-cpu-baseline --steps 10 > gpurun_out/r2p_bench_full.json 2> gpurun_out/r2p_bench_full.err
python tools/roofline_table.py gpurun_out/r2p_step_detail.tsv > gpurun_out/r2p_per_layer_roofline.md; tail -5 gpurun_out/r2p_per_layer_roofline.md
grep -n "head\|first" gpurun_out/r2p_step_detail.tsv | cut -c1-120

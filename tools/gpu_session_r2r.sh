#!/bin/bash
# x-shifted weight-gradient form for 64-channel blocks: kernel tests, microbenchmarks with / without, whole tests, step A/B
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "wgrad" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2r_tests_wgrad.log; grep -n "^E  \|passed\|failed" gpurun_out/r2r_tests_wgrad.log | cut -c1-300 | head
out=gpurun_out/r2r_wgrad.txt; rm -f $out
for s in "128 256 256 64 64" "128 256 256 128 64" "128 128 128 64 128" "128 128 128 128 128"; do
  python tools/profile_layer.py wgrad $s 5 >> $out 2>&1
  ONET_WG_NO_XSHIFT=1 python tools/profile_layer.py wgrad $s 5 2>&1 | sed 's/^/noxs /' >> $out
done
cat $out
python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2r_tests.log; grep -n "^E  \|passed\|failed" gpurun_out/r2r_tests.log | cut -c1-300 | head
for v in xs noxs xs noxs; do
  if [ $v = noxs ]; then export ONET_WG_NO_XSHIFT=1; else unset ONET_WG_NO_XSHIFT; fi
  python bench.py --no-extra --no-profile --no-cpu-baseline --steps 30 > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err
  python - $v <<'PY'
import json, sys
v=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2r_bench_{v}.json').read().strip().splitlines()[-1])
print(v, "value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY
done

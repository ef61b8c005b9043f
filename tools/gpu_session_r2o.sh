#!/bin/bash
# elementwise kernels after the multiply-high index decomposition: microbenchmarks, tests, bench
mkdir -p gpurun_out
out=gpurun_out/r2o_elementwise.txt; rm -f $out
for s in "bnapply 128 256 256 64 64" "bnapply_pool 128 256 256 64 64" "bnapply 128 128 128 128 128" "bnapply_pool 128 128 128 128 128" "bnapply 128 64 64 256 256" \
         "bnbwd_pool 128 256 256 64 64" "bnbwd_pool_g2 128 256 256 64 64" "bnbwd_pool 128 128 128 128 128" "bnbwd_pool 128 64 64 256 256" "bnbwd 128 256 256 64 64"; do
  python tools/profile_layer.py $s 5 >> $out 2>&1
done
cat $out
python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2o_tests.log; grep -n "^E  \|passed\|failed" gpurun_out/r2o_tests.log | cut -c1-300 | head
python bench.py --no-extra --no-profile --steps 30 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print("value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY

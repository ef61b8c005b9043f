#!/bin/bash
mkdir -p gpurun_out
for v in "" "ONET_NO_WGRAD_OVERLAP=1" "" "ONET_NO_WGRAD_OVERLAP=1"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-profile > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
  python - "$v" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print(f"[{sys.argv[1]:24s}] value {d['value']:.1f} ms {d['ms_per_step']:.3f} e2e {d['e2e']['value']:.1f} clocks {d['clocks']['sm_mhz']} {d['clocks']['power_w_median']}")
PY
done

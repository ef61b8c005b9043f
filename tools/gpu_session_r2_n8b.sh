#!/bin/bash
# 8 GPUs, final code of round 2: the driver's N=8 training line (no sub-records), exactly as the driver launches it
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
echo "n8 rc=$?"; tail -2 gpurun_out/r2_bench_n8_final.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n8_final.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks'])
PY

#!/bin/bash
# 2 GPUs, BASELINE configs[2] as written: global batch 512 = 256 frames per GPU (no per-kernel pass: a second activation set would not fit)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 420 $RUN --master-port 29525 bench.py --gpus 2 --steps 5 --warmup 3 --batch 256 --no-extra --no-profile > gpurun_out/r2_bench_n2_b256.json 2> gpurun_out/r2_bench_n2_b256.err
echo "n2 b256 rc=$?"; tail -2 gpurun_out/r2_bench_n2_b256.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n2_b256.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'mem', d.get('peak_memory_gb'))
PY

#!/bin/bash
# 4 GPUs: configs[2] at global batch 512 = 128 frames per GPU, and the driver's own N=4 line (64 per GPU)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29531 bench.py --gpus 4 --steps 8 --warmup 3 --batch 128 --no-extra > gpurun_out/r2_bench_n4_b128.json 2> gpurun_out/r2_bench_n4_b128.err
echo "n4 b128 rc=$?"; tail -2 gpurun_out/r2_bench_n4_b128.err | cut -c1-300
timeout 600 $RUN --master-port 29532 bench.py --gpus 4 --steps 10 --warmup 3 --no-extra > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
echo "n4 rc=$?"; tail -2 gpurun_out/r2_bench_n4.err | cut -c1-300
python - <<'PY'
import json
for f in ('r2_bench_n4_b128', 'r2_bench_n4'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'mem', d.get('peak_memory_gb'))
    except Exception as e:
        print(f, 'unreadable', e)
PY

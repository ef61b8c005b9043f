#!/bin/bash
# BASELINE configs[3] as written: the ZY-3 patch shape (3 x 224 x 224, 64 patches per GPU) on 8 GPUs, final code of round 2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29541 bench.py --gpus 8 --workload zy3 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_zy3_n8_final.json 2> gpurun_out/r2_bench_zy3_n8_final.err
echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_zy3_n8_final.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks'])
PY

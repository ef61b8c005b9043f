#!/bin/bash
# N GPUs (argument), final code of round 2: the driver's training line (no sub-records), exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err
echo "n$N rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f'gpurun_out/r2_bench_n{n}_final.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks'])
PY

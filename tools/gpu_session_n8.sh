#!/bin/bash
# 8-GPU check: bench exactly as the driver launches it, then tiled inference over 8 ranks (BASELINE configs[2], configs[4])
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 rc=$?"; wc -l gpurun_out/bench_n8.json; tail -2 gpurun_out/bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1]); print('N8', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/bench_infer.py --frames 8 --tile 2048 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/infer_n8.json 2> gpurun_out/infer_n8.err
echo "infer n8 rc=$?"; tail -1 gpurun_out/infer_n8.json | cut -c1-400

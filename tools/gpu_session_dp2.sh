#!/bin/bash
# 2-GPU session: the bench at N=2 exactly as the driver launches it (graph replay + NCCL), clean exit, then tiled inference
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 (graph) rc=$?"; wc -l gpurun_out/bench_n2.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('N2 graph', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "bench ref n2 rc=$?"; wc -l gpurun_out/bench_ref_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/bench_infer.py --frames 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/infer_n2.json 2> gpurun_out/infer_n2.err
echo "infer n2 rc=$?"; tail -1 gpurun_out/infer_n2.json | cut -c1-300

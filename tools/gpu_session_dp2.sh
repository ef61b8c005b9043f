#!/bin/bash
# 2-GPU session: DP parity on NCCL + the bench at N=2 exactly as the driver launches it
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py > gpurun_out/dp_parity.log 2>&1
echo "dp_parity rc=$?"; grep -E "dp_parity|Error|error" gpurun_out/dp_parity.log | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 (graph) rc=$?"; tail -c 600 gpurun_out/bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('N2 graph', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/bench_n2_eager.json 2> gpurun_out/bench_n2_eager.err
echo "bench n2 (eager) rc=$?"; tail -c 300 gpurun_out/bench_n2_eager.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_eager.json')); print('N2 eager', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"

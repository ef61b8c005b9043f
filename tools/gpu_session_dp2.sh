#!/bin/bash
# 2-GPU session: DP parity on NCCL + the bench at N=2 exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py > gpurun_out/dp_parity.log 2>&1
echo "dp_parity rc=$?"; tail -5 gpurun_out/dp_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json

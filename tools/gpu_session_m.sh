#!/bin/bash
# 2-GPU re-verification after the two-stream backward: DP parity, bench as the driver launches it, synth tests
mkdir -p gpurun_out
python -m pytest tests/test_synth_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py > gpurun_out/dp_parity_v4.log 2>&1
echo "dp_parity rc=$?"; grep dp_parity gpurun_out/dp_parity_v4.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_v4.json 2> gpurun_out/bench_n2_v4.err
echo "bench n2 (graph) rc=$?"; wc -l gpurun_out/bench_n2_v4.json; tail -3 gpurun_out/bench_n2_v4.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_v4.json').read().strip().splitlines()[-1]); print('N2 graph', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks'])"

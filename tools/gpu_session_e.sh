#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/profile_layer.py convT 128 128 128 128 64 5
python tools/profile_layer.py convT 128 64 64 256 128 5
python tools/profile_layer.py convT 128 32 32 512 256 5
python tools/profile_layer.py convT 128 16 16 1024 512 5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
echo "bench graph rc=$?"; tail -3 gpurun_out/bench_i.err
python -c "
import json; d=json.load(open('gpurun_out/bench_i.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks']); print(sum(v['ms_per_step'] for v in d['kernels'].values())); [print(k, v) for k,v in d['kernels'].items()]"

#!/bin/bash
# GPU session D: tests, bench (graph vs eager), bandwidth sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
ONET_BENCH_DETAIL=gpurun_out/detail.tsv python bench.py --steps 10 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
echo "bench graph rc=$?"; tail -3 gpurun_out/bench_h.err
python -c "
import json; d=json.load(open('gpurun_out/bench_h.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks']); print(sum(v['ms_per_step'] for v in d['kernels'].values())); [print(k, v) for k,v in d['kernels'].items()]"
python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_h_eager.json 2> gpurun_out/bench_h_eager.err
echo "bench eager rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_h_eager.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"
python tools/profile_layer.py first_fwd 128 256 256 1 64 5
python tools/profile_layer.py first_wgrad 128 256 256 1 64 5
python tools/profile_layer.py first_fwd 128 224 224 3 64 5
python tools/profile_layer.py first_wgrad 128 224 224 3 64 5

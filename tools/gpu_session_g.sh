#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ONET_BENCH_DETAIL=gpurun_out/detail.tsv python bench.py --steps 10 --warmup 3 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
echo "bench graph rc=$?"; tail -3 gpurun_out/bench_j.err
python -c "
import json; d=json.load(open('gpurun_out/bench_j.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks']); print(sum(v['ms_per_step'] for v in d['kernels'].values())); [print(k, v) for k,v in d['kernels'].items()]"

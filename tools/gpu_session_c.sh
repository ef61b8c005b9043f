#!/bin/bash
# GPU session C: tests, bench (graph vs eager)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
echo "bench graph rc=$?"; tail -3 gpurun_out/bench_g.err
python -c "
import json; d=json.load(open('gpurun_out/bench_g.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks']); print(sum(v['ms_per_step'] for v in d['kernels'].values()))"
python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_g_eager.json 2> gpurun_out/bench_g_eager.err
echo "bench eager rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_g_eager.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"

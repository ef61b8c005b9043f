#!/bin/bash
# tests (single process) + single-layer timings + bench with per-call detail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "fwd 128 64 64 256 256" "fwd 128 256 256 64 64" "fwd 128 256 256 128 64" "fwd 128 128 128 128 128" "wgrad 128 64 64 256 256" "wgrad 128 256 256 64 64" "wgrad 128 128 128 128 128" "wgrad 128 16 16 1024 1024"; do
  python tools/profile_layer.py $cfg 5 2>&1 | tail -1
done
ONET_BENCH_DETAIL=gpurun_out/detail.tsv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_d.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved']); [print(k, v) for k,v in d['kernels'].items()]"

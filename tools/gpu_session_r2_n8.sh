#!/bin/bash
# 8 GPUs: the driver's own N=8 line with its sub-records (inference configs[4], ZY-3 configs[3]), the all-reduce timeline,
# and one A/B of the NCCL CTA budget
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "n8 rc=$?"; tail -2 gpurun_out/r2_bench_n8.err | cut -c1-300
timeout 300 $RUN --master-port 29542 tools/overlap_timeline.py > gpurun_out/r2_overlap_n8.txt 2> gpurun_out/r2_overlap_n8.err
echo "timeline rc=$?"; tail -3 gpurun_out/r2_overlap_n8.txt
NCCL_MAX_CTAS=8 timeout 600 $RUN --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/r2_bench_n8_maxctas8.json 2> gpurun_out/r2_bench_n8_maxctas8.err
echo "n8 maxctas8 rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_n8', 'r2_bench_n8_maxctas8'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1),
              {k: (round(v['value'], 1), round(v['e2e']['value'], 1)) for k, v in d.items() if isinstance(v, dict) and 'e2e' in v and k in ('infer', 'zy3')})
    except Exception as e:
        print(f, 'unreadable', e)
PY

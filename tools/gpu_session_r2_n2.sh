#!/bin/bash
# 2 GPUs: configs[2] at global batch 512 = 256 frames per GPU (peak memory), the driver's own N=2 line, overlap timeline, DP parity
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 --batch 256 --no-extra > gpurun_out/r2_bench_n2_b256.json 2> gpurun_out/r2_bench_n2_b256.err
echo "n2 b256 rc=$?"; tail -2 gpurun_out/r2_bench_n2_b256.err | cut -c1-300
timeout 600 $RUN --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "n2 rc=$?"; tail -2 gpurun_out/r2_bench_n2.err | cut -c1-300
timeout 300 $RUN --master-port 29523 tools/overlap_timeline.py > gpurun_out/r2_overlap_n2.txt 2> gpurun_out/r2_overlap_n2.err
echo "timeline rc=$?"; tail -4 gpurun_out/r2_overlap_n2.txt; tail -2 gpurun_out/r2_overlap_n2.err | cut -c1-300
timeout 300 $RUN --master-port 29524 tools/dp_parity.py > gpurun_out/r2_dp_parity_n2.log 2>&1
echo "dp_parity rc=$?"; grep dp_parity gpurun_out/r2_dp_parity_n2.log | tail -4
python - <<'PY'
import json
for f in ('r2_bench_n2_b256', 'r2_bench_n2'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'mem', d.get('peak_memory_gb'),
              {k: (round(v['value'], 1), round(v['e2e']['value'], 1)) for k, v in d.items() if isinstance(v, dict) and 'e2e' in v and k in ('infer', 'zy3')})
    except Exception as e:
        print(f, 'unreadable', e)
PY

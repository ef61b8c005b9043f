#!/bin/bash
# one GPU session: bench with per-call detail, ncu launch list of one training step, ncu full captures of single layers
mkdir -p gpurun_out
ONET_BENCH_DETAIL=gpurun_out/detail.tsv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 570 -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
for cfg in "fwd 128 64 64 256 256" "fwd 128 256 256 64 64" "wgrad 128 64 64 256 256" "wgrad 128 256 256 64 64"; do
  tag=$(echo $cfg | tr ' ' '_')
  python tools/profile_layer.py $cfg 3 > gpurun_out/layer_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:tapgemm -s 2 -c 1 -o gpurun_out/prof_$tag \
      python tools/profile_layer.py $cfg 1 > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"; cat gpurun_out/layer_$tag.log
done

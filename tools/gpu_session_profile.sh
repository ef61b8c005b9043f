#!/bin/bash
# Profile session: bench (detail), ncu launch list of the eager step, ncu --set full of the dominant kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ONET_BENCH_DETAIL=gpurun_out/detail.tsv python bench.py --steps 10 --warmup 3 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err
echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 650 -c 240 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
prof() {  # prof <skip> <count> <kind> <N> <H> <W> <Cin> <Cout>
  local skip=$1 count=$2; shift 2
  local tag=$(echo "$@" | tr ' ' '_')
  python tools/profile_layer.py "$@" 3 > gpurun_out/layer_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'halo|tapgemm|bn_|conv_first|wgrad' -s $skip -c $count -o gpurun_out/prof_$tag -f \
      python tools/profile_layer.py "$@" 1 > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"; cat gpurun_out/layer_$tag.log
}
prof 2 1 fwd 128 64 64 256 256
prof 2 1 fwd 128 128 128 128 128
prof 2 1 dgrad 128 256 256 64 64
prof 2 1 wgrad 128 64 64 256 256
prof 2 1 wgrad 128 256 256 64 64
prof 6 2 bnbwd 128 256 256 64 64
prof 2 1 bnapply_pool 128 256 256 64 64

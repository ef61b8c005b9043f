#!/bin/bash
# round-1 v4 evidence: all GPU tests, bench line, PyTorch-eager comparator on the same box, inference, ncu launch list + BN-bwd captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ONET_BENCH_DETAIL=gpurun_out/detail_v4.tsv python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_v4.err
python tests/compare_torch_gpu.py --batch 64 --steps 5 > gpurun_out/torch_gpu.jsonl 2> gpurun_out/torch_gpu.err; echo "torch comparator rc=$?"; cat gpurun_out/torch_gpu.jsonl
python tools/bench_infer.py --frames 4 --tile 2048 --steps 5 --no-cpu-baseline > gpurun_out/infer_n1_tile2048.json 2> gpurun_out/infer_n1_tile2048.err; echo "infer2048 rc=$?"; cat gpurun_out/infer_n1_tile2048.json
python tools/bench_infer.py --frames 4 --tile 1024 --steps 5 > gpurun_out/infer_n1_tile1024.json 2> gpurun_out/infer_n1_tile1024.err; echo "infer1024 rc=$?"; cat gpurun_out/infer_n1_tile1024.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 650 -c 240 --csv --log-file gpurun_out/launches_v4.csv \
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
prof 6 2 bnbwd_pool_g2 128 256 256 64 64
prof 6 2 bnbwd_pool 128 128 128 128 128

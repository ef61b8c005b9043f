#!/bin/bash
# GPU session B: tests, bench with per-call detail, bandwidth-kernel sweep, ncu full captures of convT / first-layer kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ONET_BENCH_DETAIL=gpurun_out/detail.tsv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved']); [print(k, v) for k,v in d['kernels'].items()]"
python tools/profile_layer.py sweep_bw 64 > gpurun_out/sweep_b.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep_b.log
prof() {  # prof <skip> <count> <kind> <N> <H> <W> <Cin> <Cout>
  local skip=$1 count=$2; shift 2
  local tag=$(echo "$@" | tr ' ' '_')
  python tools/profile_layer.py "$@" 3 > gpurun_out/layer_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'halo|tapgemm|bn_|conv_first|colsum' -s $skip -c $count -o gpurun_out/prof_$tag -f \
      python tools/profile_layer.py "$@" 1 > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"; cat gpurun_out/layer_$tag.log
}
prof 2 1 convT 128 128 128 128 64
prof 2 1 convT_wgrad 128 128 128 128 64
prof 2 1 convT_dgrad 128 128 128 128 64
prof 2 1 first_fwd 128 256 256 1 64
prof 2 1 first_wgrad 128 256 256 1 64
prof 6 2 bnbwd_pool 128 256 256 64 64

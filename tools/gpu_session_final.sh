#!/bin/bash
# round-1 final evidence (v5): all GPU tests, smoke, bench line (both arms), ncu launch list of the eager step, ncu --set full of the
# fused dgrad + BatchNorm-reduce kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
ONET_BENCH_DETAIL=gpurun_out/detail_v5.tsv python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_v5.err
python tools/roofline_table.py gpurun_out/detail_v5.tsv > gpurun_out/per_layer_roofline_v5.md; tail -5 gpurun_out/per_layer_roofline_v5.md
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v5.json 2> gpurun_out/bench_ref_v5.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_v5.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 650 -c 240 --csv --log-file gpurun_out/launches_v5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
cat > /tmp/red_probe.py <<'PY'
import sys, os, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import gpu_util as U
n, h, w, c = 128, 64, 64, 256
bf = torch.bfloat16
dy = (torch.randn(n, h, w, c, device="cuda") * 0.5).to(bf)
wt = torch.randn(c, c, 3, 3, device="cuda") * 0.01
_, wd = U.pack_conv(wt, U.BF16)
y = (torch.randn(n, h, w, c, device="cuda") * 1.5 + 0.3).to(bf)
aff = torch.rand(4, 2, c, device="cuda") + 0.5
out = torch.empty(n, h, w, c, dtype=bf, device="cuda")
sums = torch.zeros(2, 2, c, dtype=torch.float64, device="cuda")
for _ in range(3):
    U.call("onet_conv3x3_dgrad_bnred", U.ptr(dy), c, 0, n, h, w, c, U.ptr(wd), c, U.ptr(out), U.ptr(y), U.ptr(aff[2]), U.ptr(aff[3]),
           U.ptr(aff[0]), U.ptr(aff[1]), U.ptr(sums), n // 2, U.BF16, U.ENGINE_TC, U.stream())
torch.cuda.synchronize()
PY
python /tmp/red_probe.py && ncu --set full --clock-control none --import-source on -k regex:'halo2' -s 2 -c 1 -o gpurun_out/prof_dgrad_bnred_128_64_64_256_256 -f python /tmp/red_probe.py > gpurun_out/ncu_dgrad_bnred.log 2>&1
echo "ncu bnred rc=$?"

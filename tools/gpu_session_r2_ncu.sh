#!/bin/bash
# ncu evidence: launch list of one eager training step, `ncu --set full` captures of the dominant kernels (each after its own
# command exited 0 without ncu)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extra > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 620 -c 260 --csv --log-file gpurun_out/r2_launches_step.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extra > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_step.csv
cap() {  # kind N H W Cin Cout kernel-regex count
  python tools/profile_layer.py $1 $2 $3 $4 $5 $6 3 && \
  ncu --set full --clock-control none --import-source on -k regex:"$7" -s $8 -c $9 -o gpurun_out/r2_prof_$1_$2_$3_$4_$5_$6 -f \
      python tools/profile_layer.py $1 $2 $3 $4 $5 $6 3 > gpurun_out/r2_ncu_$1.log 2>&1
  echo "capture $1 rc=$?"
  # summarise on the box (the reports are ~16 MB each; gpurun brings back at most 64 MiB): text summary + DRAM traffic JSON
  python tools/ncu_summary.py gpurun_out/r2_prof_$1_$2_$3_$4_$5_$6.ncu-rep >> gpurun_out/r2_ncu_summary.txt
  python tools/ncu_traffic.py gpurun_out/r2_prof_$1_$2_$3_$4_$5_$6.ncu-rep > gpurun_out/r2_traffic_$1_$5_$6.json
  if [ "$1$5$6" != "fwd256256" ]; then rm -f gpurun_out/r2_prof_$1_$2_$3_$4_$5_$6.ncu-rep; fi
}
rm -f gpurun_out/r2_ncu_summary.txt gpurun_out/r2_traffic_*.json
cap fwd 128 64 64 256 256 'halo2_px' 2 1
cap wgrad 128 64 64 256 256 'wgrad3x3_halo2' 2 1
cap dgrad 128 256 256 64 64 'halo_res_px' 2 1
cap fwd 128 256 256 64 64 'halo_res_px' 2 1
cap wgrad 128 256 256 64 64 'wgrad3x3_halo_kernel' 2 1
cap bnbwd 128 256 256 64 64 'bn_bwd_px' 4 2
cap bnbwd_pool 128 128 128 128 128 'bn_bwd_win' 4 2
cap bnapply_pool 128 256 256 64 64 'bn_relu_apply' 2 1
cap fwd_tf32 128 64 64 256 256 'halo2_px' 2 1
cap wgrad_tf32 128 64 64 256 256 'tapgemm_wg' 2 1
# first conv on warp-level MMAs (driver: tools/bench_first_layer.py)
python tools/bench_first_layer.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'first_mma' -s 6 -c 2 -o gpurun_out/r2_prof_first_mma -f python tools/bench_first_layer.py > gpurun_out/r2_ncu_first.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_prof_first_mma.ncu-rep >> gpurun_out/r2_ncu_summary.txt
rm -f gpurun_out/r2_prof_first_mma.ncu-rep
python - <<'PY'
import glob, json
out = []
for f in sorted(glob.glob('gpurun_out/r2_traffic_*.json')):
    out += json.load(open(f))
json.dump(out, open('gpurun_out/r2_ncu_traffic.json', 'w'), indent=1)
for c in out:
    print(c['capture'], c['kernel'][:50], c['duration_us'], 'us dram', c['dram_bytes'], 'alg', c['algorithmic_bytes'], 'tc%', c['tcgen05_pct_of_peak'], 'dram%', c['dram_pct_of_peak'])
PY
du -sh gpurun_out

#!/bin/bash
# why does the statistics epilogue of the 64-column tiles cost 14 %?  ncu --set full of the same layer with and without statistics
mkdir -p gpurun_out
for k in fwd dgrad; do
  ncu --set full --clock-control none -k regex:'halo_res_px' -s 2 -c 1 -o gpurun_out/r2n_$k -f python tools/profile_layer.py $k 128 256 256 64 64 3 > gpurun_out/r2n_ncu_$k.log 2>&1
  python tools/ncu_stalls.py gpurun_out/r2n_$k.ncu-rep > gpurun_out/r2n_stalls_$k.txt
  rm -f gpurun_out/r2n_$k.ncu-rep
done
wc -l gpurun_out/r2n_stalls_*.txt

#!/bin/bash
# where do the epilogue instructions of the 64-column forward conv go?  source-level counts of one launch
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'halo_res_px' -s 2 -c 1 -o gpurun_out/r2n_fwd -f python tools/profile_layer.py fwd 128 256 256 64 64 3 > gpurun_out/r2n_ncu_fwd.log 2>&1
ncu -i gpurun_out/r2n_fwd.ncu-rep --page source --csv > gpurun_out/r2n_source_raw.csv 2> gpurun_out/r2n_source_raw.err
head -c 1500 gpurun_out/r2n_source_raw.csv
python tools/ncu_source.py gpurun_out/r2n_fwd.ncu-rep 60 > gpurun_out/r2n_source_fwd.txt 2>&1
head -30 gpurun_out/r2n_source_fwd.txt
gzip -f gpurun_out/r2n_source_raw.csv
rm -f gpurun_out/r2n_fwd.ncu-rep

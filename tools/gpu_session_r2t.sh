#!/bin/bash
# sensitivity of the BatchNorm-backward kernels to the grid size
mkdir -p gpurun_out
out=gpurun_out/r2t_bnbwd_grid.txt; rm -f $out
for b in 296 592 1184 2368; do
  for s in "bnbwd 128 256 256 64 64" "bnbwd 128 128 128 128 128" "bnbwd_pool 128 256 256 64 64" "bnbwd_pool 128 128 128 128 128"; do
    echo -n "blocks=$b " >> $out
    ONET_BN_BWD_BLOCKS=$b python tools/profile_layer.py $s 5 >> $out 2>&1
  done
done
cat $out

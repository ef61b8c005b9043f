#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "conv3x3_tc" 2>&1 | tail -4
for cfg in "wgrad 128 128 128 128 128" "wgrad 128 128 128 256 128" "wgrad 128 64 64 256 256" "wgrad 128 64 64 512 256" "wgrad 128 32 32 512 512" "wgrad 128 32 32 1024 512" "wgrad 128 16 16 1024 1024" "wgrad 128 64 64 128 256"; do
  timeout 120 python tools/profile_layer.py $cfg 5 2>&1 | tail -1
  ONET_NO_2CTA_WGRAD=1 timeout 120 python tools/profile_layer.py $cfg 5 2>&1 | tail -1 | sed 's/^/   (1-CTA) /'
done

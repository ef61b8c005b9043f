#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "conv3x3_tc" 2>&1 | tail -4
ONET_2CTA_MIN_KC=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "conv3x3_tc" 2>&1 | tail -2
for cfg in "dgrad 128 256 256 64 64" "fwd 128 256 256 64 64" "dgrad 128 256 256 64 128" "fwd 128 128 128 64 128" "fwd 128 128 128 128 128" "fwd 128 256 256 128 64" "dgrad 128 256 256 128 64" "fwd 128 64 64 128 256" "dgrad 128 128 128 128 64" "fwd 128 16 16 1024 1024" "fwd 128 16 16 512 1024" "fwd 128 32 32 256 512"; do
  ONET_2CTA_MIN_KC=1 timeout 120 python tools/profile_layer.py $cfg 5 2>&1 | tail -1
  ONET_NO_2CTA=1 timeout 120 python tools/profile_layer.py $cfg 5 2>&1 | tail -1 | sed 's/^/   (1-CTA) /'
done

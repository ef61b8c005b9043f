#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_training_trajectory_gpu.py -m gpu -q -x -s 2>&1 | grep -v Warning | tail -8
python tools/bench_infer.py --frames 4 --tile 2048 --steps 5 > gpurun_out/infer_n1_tile2048_v5.json 2> gpurun_out/infer_v5.err; echo "infer rc=$?"; cat gpurun_out/infer_n1_tile2048_v5.json | cut -c1-900

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/bench_infer.py --frames 2 --steps 3 --warmup 2 > gpurun_out/infer_n1.json 2> gpurun_out/infer_n1.err; echo "infer rc=$?"; tail -3 gpurun_out/infer_n1.err; cut -c1-330 gpurun_out/infer_n1.json

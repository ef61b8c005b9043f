#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_synth_gpu.py -m gpu -q -x 2>&1 | tail -12
python - <<'PY'
import time, torch
from onet_b200 import synth
synth.k_correlated_background(2, 400, seed=1); torch.cuda.synchronize()
t0=time.perf_counter(); a=synth.k_correlated_background(64, 400, seed=2); torch.cuda.synchronize(); dt=time.perf_counter()-t0
print(f"correlated K field: 64 frames of 400x400 in {dt*1e3:.1f} ms = {64/dt:.0f} frames/s; E[a^2]={float((a**2).mean()):.3f}")
PY

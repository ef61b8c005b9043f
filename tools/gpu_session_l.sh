#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_synth_gpu.py -m gpu -q -x 2>&1 | tail -15
python - <<'PY'
import time, torch
from onet_b200 import synth
for kind, fn in (("rayleigh", synth.get_rayleigh_frames), ("kdist", synth.get_k_frames)):
    fn(8, snr=5); torch.cuda.synchronize()
    t0 = time.perf_counter(); f, m = fn(512, snr=5, seed=3); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{kind}: 512 frames of 400x400 with 20 targets in {dt*1e3:.1f} ms = {512/dt:.0f} frames/s (host target table included)")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bg = (synth.rayleigh_background if kind == "rayleigh" else synth.k_background)(512, 400, 400, seed=4); e1.record(); torch.cuda.synchronize()
    print(f"   background kernel alone: {e0.elapsed_time(e1):.3f} ms for {bg.numel()*4/1e6:.0f} MB = {bg.numel()*4/e0.elapsed_time(e1)/1e6:.0f} GB/s written")
PY

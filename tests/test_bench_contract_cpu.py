"""bench.py contract without a GPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys the driver reads,
rank != 0 under a multi-rank launch prints nothing and exits 0, and the GPU arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=timeout,
                          cwd=ROOT)


@pytest.mark.timeout(900)
def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "onet_train_images_per_sec" and d["unit"] == "images/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "images/s" and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"},
             timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the GPU arm would run")
    r = _run(["--steps", "1", "--warmup", "3"], timeout=300)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_roofline_table_tool_runs_on_the_committed_step_dump():
    """tools/roofline_table.py runs over the committed per-call dump (peaks from MEASURED_PEAKS.json or the recipe's fallback)."""
    dump = os.path.join(ROOT, "profiles", "r1_step_detail_v5.tsv")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "roofline_table.py"), dump], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert "| call | kernel | ms |" in r.stdout and "Step (serial sum of the calls)" in r.stdout
    total = [ln for ln in r.stdout.splitlines() if ln.startswith("**Step")][0]
    frac = float(total.split("=")[-1].strip(" .*"))
    assert 0.5 < frac < 1.0, total          # the committed table (profiles/r1_per_layer_roofline_v5.md) says 0.83 with this pod's peaks

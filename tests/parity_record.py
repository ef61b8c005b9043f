"""Measured parity values of the -m gpu tests, kept: every test that compares the CUDA path with the oracle calls
`record(key, **values)`; the values are merged into `gpurun_out/r2_parity.json` (the directory gpurun brings back from the GPU
box; copied to `profiles/r2_parity.json` afterwards) so that the numbers behind the assertions are on file, not just dots."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "gpurun_out", "r2_parity.json")


def record(key, **values):
    try:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        data = {}
        if os.path.isfile(PATH):
            with open(PATH) as f:
                data = json.load(f)
        data[key] = {k: (float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v) for k, v in values.items()}
        with open(PATH, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass            # a read-only checkout must not fail a parity test

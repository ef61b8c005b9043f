"""Aggregate the source page of an .ncu-rep (ncu --set full --import-source on, -lineinfo build) by CUDA source line:
    python tools/ncu_source.py rep.ncu-rep [top]
Prints the columns found, then the top lines by warp-stall samples and by executed instructions."""
import csv
import subprocess
import sys
from collections import defaultdict


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if any("Instructions Executed" in c for c in r))
    hdr = rows[hdr_i]
    print("columns:", hdr)
    ci = {h: i for i, h in enumerate(hdr)}
    c_src = next((ci[h] for h in hdr if h.strip() in ("Source", "CUDA Source", "#")), 0)
    c_exec = next(ci[h] for h in hdr if h.strip() == "Instructions Executed")
    c_samp = next((ci[h] for h in hdr if "Sampling Data (All)" in h), None)
    agg = defaultdict(lambda: [0.0, 0.0, 0])
    for r in rows[hdr_i + 1:]:
        if len(r) <= c_exec:
            continue
        key = r[c_src][:150]
        a = agg[key]
        a[0] += num(r[c_samp]) if c_samp is not None else 0.0
        a[1] += num(r[c_exec])
        a[2] += 1
    tot_s = sum(a[0] for a in agg.values()) or 1.0
    tot_e = sum(a[1] for a in agg.values()) or 1.0
    print(f"total samples {tot_s:.0f}, total warp instructions {tot_e:.0f}, {len(agg)} distinct source keys")
    print("--- by stall samples")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * a[0] / tot_s:6.2f}% samp {100 * a[1] / tot_e:6.2f}% inst  n={a[2]:4d}  {k}")
    print("--- by executed instructions")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * a[1] / tot_e:6.2f}% inst {100 * a[0] / tot_s:6.2f}% samp  n={a[2]:4d}  {k}")


if __name__ == "__main__":
    main()

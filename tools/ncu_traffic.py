"""DRAM traffic of captured launches vs their algorithmic bytes, from `.ncu-rep` files (ncu --set full):

    python tools/ncu_traffic.py gpurun_out/r2_prof_*.ncu-rep > profiles/r2_ncu_traffic.json

The capture name encodes the launch: r2_prof_<kind>_<N>_<H>_<W>_<Cin>_<Cout>.ncu-rep (kinds of tools/profile_layer.py).  Output:
a JSON list with, per captured kernel launch, duration, dram__bytes_read + dram__bytes_write, the algorithmic bytes of that launch
(every operand read once, the result written once), tcgen05 and DRAM utilisation.  bench.py attaches the matching entries to its
`roofline.traffic` / `roofline_hbm.traffic` keys (`family` = the C-ABI call the kernel belongs to)."""
import csv
import json
import os
import subprocess
import sys


def algorithmic(kind, kernel, n, h, w, ci, co):
    e = 4.0 if kind.endswith("tf32") else 2.0
    px = n * h * w
    if kind.startswith(("fwd", "dgrad")):
        return px * (ci + co) * e + 9 * ci * co * e, "conv3x3_fwd"
    if kind.startswith("wgrad"):
        return px * (ci + co) * e + 9 * ci * co * 4, "conv3x3_wgrad"
    if kind.startswith("bnapply"):
        return px * co * (2 * e + (e / 4 if "pool" in kind else 0)), "bn_relu_apply"
    if kind.startswith("bnbwd"):
        reads = 2 * e + (e if kind.endswith("_g2") else 0) + (e / 4 if "pool" in kind else 0)
        apply_pass = kernel.rstrip(">").rstrip().endswith(("1", "true"))        # template argument APPLY
        return px * co * (reads + (e if apply_pass else 0)), "bn_relu_bwd"
    if kind.startswith("convT"):
        return px * (ci + 4 * (ci // 2)) * e, "convT2x2"
    return 0.0, kind


def main():
    out = []
    for f in sys.argv[1:]:
        name = os.path.basename(f).replace(".ncu-rep", "")
        parts = name.split("_")
        dims = [int(v) for v in parts[-5:]]
        kind = "_".join(parts[2:-5]) if parts[0] == "r2" else "_".join(parts[1:-5])
        raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        idx = {h: i for i, h in enumerate(rows[0])}
        units = rows[1]

        def val(r, key, scale_bytes=False):
            if key not in idx or r[idx[key]] == "":
                return None
            v = float(r[idx[key]].replace(",", ""))
            if scale_bytes:
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[key]], 1)
            return v
        for r in rows[2:]:
            kern = r[idx["Kernel Name"]].replace("onet::", "").split("(")[0].replace("void ", "")
            dur = val(r, "gpu__time_duration.sum")
            dur_us = dur * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1)
            rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
            alg, fam = algorithmic(kind, kern, *dims)
            out.append(dict(capture=f"{kind}_" + "_".join(str(d) for d in dims), kernel=kern, family=fam, duration_us=round(dur_us, 1),
                            dram_bytes=int((rd or 0) + (wr or 0)), algorithmic_bytes=int(alg),
                            tcgen05_pct_of_peak=val(r, "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed")
                            or val(r, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                            dram_pct_of_peak=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

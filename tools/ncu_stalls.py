"""Print issue / stall metrics of the launches in .ncu-rep files (ncu --set full):  python tools/ncu_stalls.py a.ncu-rep [...]"""
import csv
import subprocess
import sys

PAT = ("issue_stalled", "inst_executed.sum", "issue_active", "cycles_elapsed.avg", "cycles_active.avg", "pipe_tensor", "utchmma",
       "inst_executed_pipe_lsu", "inst_executed_pipe_alu", "inst_executed_pipe_fma", "data_pipe", "shared", "warps_eligible",
       "gpu__time_duration.sum", "launch__registers")


def main():
    for f in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print(f"== {f}: no data")
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            print(f"== {f} :: {kn[:90]}")
            for h, u, v in zip(hdr, units, vals):
                if any(p in h for p in PAT):
                    try:
                        if float(v.replace(",", "")) == 0.0:
                            continue
                    except ValueError:
                        pass
                    print(f"   {h:100s} {v} {u}")


if __name__ == "__main__":
    main()

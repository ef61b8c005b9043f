"""Summarise .ncu-rep captures (read on the CPU box with `ncu -i`) into small text files for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [...] > profiles/rN_ncu_summary.txt
For each captured launch: duration, DRAM bytes, tensor-pipe utilisation, shared-memory port wavefronts
(tensor-core operand reads + TMA fills), L2 hit rate, registers, grid."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "sm cycles elapsed (avg)"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tcgen05 bf16 ops % of peak"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe cycles active %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smem wavefronts read by tensor core (128 B each)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes (TMA fills + loads)"),
    ("l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "smem bank reads % of peak"),
    ("l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smem bank writes % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
]


def main():
    for f in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print(f"== {f}: no data")
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            print(f"== {f.split('/')[-1]} :: {r[idx['Kernel Name']]}")
            for k, label in KEYS:
                if k in idx and r[idx[k]] != "":
                    print(f"   {label:52s} {r[idx[k]]} {units[idx[k]]}")
            print()


if __name__ == "__main__":
    main()

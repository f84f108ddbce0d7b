#!/usr/bin/env python
"""Condense an Nsight Compute report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the few
per-launch numbers the roofline discussion uses.  Usage: python profiles/summarize_ncu.py REP [REP...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/LSU data-pipe % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank-conflict wavefronts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        name_i = hdr.index("Kernel Name")
        print(f"## {rep}")
        for r in data:
            print(f"### {r[name_i][:100]}")
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"- {label}: {r[i]} {units[i]}  (`{key}`)")
            print()


if __name__ == "__main__":
    main()

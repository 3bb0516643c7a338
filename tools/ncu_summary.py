#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python tools_ncu_summary.py gpurun_out/prof.ncu-rep [regex ...]"""
import csv, io, re, subprocess, sys

KEYS = [r"gpu__time_duration.sum", r"dram__bytes_(read|write).sum$", r"gpu__dram_throughput.avg.pct", r"launch__registers_per_thread",
        r"launch__occupancy_limit", r"sm__warps_active.avg.pct_of_peak", r"l1tex__throughput.avg.pct", r"lts__throughput.avg.pct",
        r"sm__throughput.avg.pct", r"l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$", r"l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum$",
        r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?.sum$", r"l1tex__t_sector_hit_rate.pct", r"lts__t_sector_hit_rate.pct",
        r"lts__t_sectors.sum$", r"sm__cycles_elapsed.max", r"smsp__inst_executed.sum$", r"l1tex__data_bank_conflicts_pipe_lsu.sum$",
        r"smsp__average_warps_issue_stalled_.*_per_issue_active", r"smsp__warp_issue_stalled_.*_per_warp_active.pct", r"sm__inst_executed_pipe_fp64",
        r"smsp__issue_active.avg.pct", r"l1tex__lsu_writeback_active", r"l1tex__data_pipe", r"l1tex__t_set_accesses_pipe_lsu_mem_global_op_ld.sum$"]

def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    pats = [re.compile(k) for k in KEYS + extra]
    for n, d in enumerate(data):
        name = d[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"=== launch {n}: {name[:100]}")
        for h, u, v in zip(hdr, units, d):
            if any(p.search(h) for p in pats):
                try:
                    if float(v.replace(",", "")) == 0.0 and "stall" in h:
                        continue
                except ValueError:
                    pass
                print(f"  {h:95s} {v} {u}")

if __name__ == "__main__":
    main()

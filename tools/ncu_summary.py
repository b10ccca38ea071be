#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i`) into the few numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for h, u in zip(hdr, units):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                v = d[h]
                try:
                    if float(v) == 0 and "stalled" in h:
                        continue
                except ValueError:
                    pass
                print("  %-92s %16s %s" % (h, v, u))


if __name__ == "__main__":
    main()

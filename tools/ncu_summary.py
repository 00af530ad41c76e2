"""Text summary of an .ncu-rep capture for profiles/: the metrics the DESIGN/bench numbers come from.
usage: ncu_summary.py report.ncu-rep "header line" > profiles/rNN_ncu_<kernel>.txt"""
import csv, subprocess, sys

WANT = ("Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active")
PREFIX = ("smsp__average_warps_issue_stalled_",)
ALSO = ("smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio")

rep, header = sys.argv[1], sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
names, units = rows[0], rows[1]
for h in header:
    print("# " + h)
for vals in rows[2:]:          # one block per captured launch
    if len(vals) != len(names):
        continue
    for i, n in enumerate(names):
        if n in WANT or n in ALSO or (n.startswith(PREFIX) and n.endswith("_per_issue_active.ratio")):
            print("%-88s %-16s %s" % (n, units[i], vals[i]))
    print()

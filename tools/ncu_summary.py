"""ncu_summary.py <report.ncu-rep> [launch index]: the fixed list of metrics the profiles/ summaries hold"""
import csv, subprocess, sys
KEYS = ["dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for li, vals in enumerate(rows[2:]):
    if len(sys.argv) > 2 and li != int(sys.argv[2]):
        continue
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("%-103s %s" % ("Kernel Name", d["Kernel Name"][1]))
    for h in sorted(d):
        if h in KEYS or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
            print("%-90s %-12s %s" % (h, d[h][0], d[h][1]))
    print()

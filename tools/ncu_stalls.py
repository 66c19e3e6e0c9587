"""All warp-stall reasons + pipe utilisation of every kernel in a capture: python tools/ncu_stalls.py file.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    d = dict(zip(hdr, row))
    print(d.get("Kernel Name", "?")[:110])
    keys = [h for h in hdr if ("issue_stalled" in h and h.endswith("per_issue_active.ratio")) or h in (
        "gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]
    for h in keys:
        v = d[h]
        try:
            if "issue_stalled" in h and float(v) < 0.05:
                continue
        except ValueError:
            pass
        print("   %-90s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), v))
    print()

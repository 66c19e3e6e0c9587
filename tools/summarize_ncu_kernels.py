"""Per-kernel table from an `ncu --set full` capture of many launches: time, DRAM bytes, achieved GB/s, L2 sectors, shared
wavefronts.  python tools/summarize_ncu_kernels.py file.ncu-rep"""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
need = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]


def conv(v, u):
    v = float(v.replace(",", "")) if v else 0.0
    u = u.lower()
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u)
    return v * scale if scale else v


agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[ix["Kernel Name"]][:70]
    a = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, l2=0.0, sh=0.0, occ=0.0, regs=0))
    a["n"] += 1
    a["t"] += conv(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
    a["rd"] += conv(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    a["wr"] += conv(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    a["l2"] += conv(r[ix["lts__t_sectors.sum"]], "")
    a["sh"] += conv(r[ix["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]], "")
    a["occ"] += conv(r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]], "")
    a["regs"] = r[ix["launch__registers_per_thread"]]
print("%-72s %4s %9s %9s %9s %9s %10s %10s %5s %4s" % ("kernel", "n", "avg us", "rd MB", "wr MB", "GB/s", "L2 MB", "smem wf", "occ%", "regs"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    n = a["n"]
    print("%-72s %4d %9.1f %9.1f %9.1f %9.0f %10.1f %10.0f %5.0f %4s" % (k, n, a["t"] / n * 1e6, a["rd"] / n / 1e6, a["wr"] / n / 1e6,
          (a["rd"] + a["wr"]) / a["t"] / 1e9, a["l2"] * 32 / n / 1e6, a["sh"] / n, a["occ"] / n, a["regs"]))

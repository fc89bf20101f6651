"""Per-kernel summary (CSV) of an ncu --set full report: duration, DRAM bytes, instructions,
occupancy, issue activity.  usage: ncu_summary.py report.ncu-rep label > out.csv"""
import csv, subprocess, sys, io
rep, label = sys.argv[1], sys.argv[2]
M = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(h)}
w = csv.writer(sys.stdout)
w.writerow(["capture", "Kernel Name"] + M)
w.writerow(["", ""] + [units[ix[m]] if m in ix else "" for m in M])
for r in rows[2:]:
    w.writerow([label, r[ix["Kernel Name"]][:90]] + [r[ix[m]] if m in ix else "" for m in M])

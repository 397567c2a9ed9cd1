"""Summarise an .ncu-rep (raw page) into a small CSV of the metrics the roofline needs."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([n for n, _ in idx])
    w.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
for r in rows[2:]:
    print({n: r[i][:60] for n, i in idx})

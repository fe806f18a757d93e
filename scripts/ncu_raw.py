"""Print selected raw metrics of every kernel in an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("----")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {units[i]:14s} {r[i]}")

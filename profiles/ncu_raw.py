"""print selected raw metrics of every kernel in an .ncu-rep:  python profiles/ncu_raw.py <rep> [metric-substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
                        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
                        "smsp__issue_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
                        "sm__pipe_tensor_cycles_active.avg.pct", "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct",
                        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:110])
    for i, h in enumerate(hdr):
        if any(w in h for w in want):
            print(f"   {h:70s} {r[i]:>16s} {units[i]}")

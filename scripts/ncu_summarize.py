"""Dev tool (CPU): turn `ncu -i X.ncu-rep --page raw --csv` into the small per-kernel table kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python scripts/ncu_summarize.py [--one-per-kernel] > profiles/x.txt
"""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__cluster_dim_x",
    "smsp__inst_executed.sum",
]
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
data = rows[hdr + 2:]
col = {n: i for i, n in enumerate(names)}
one = "--one-per-kernel" in sys.argv
seen = {}
out = []
for r in data:
    if len(r) < len(names):
        continue
    k = r[col["Kernel Name"]]
    short = k.split("(")[0].replace("buddy::", "")
    seen[short] = seen.get(short, 0) + 1
    if one and seen[short] > 1:
        continue
    out.append((short + (f"#{seen[short]}" if not one else ""), r))
print("# columns: " + " | ".join(n for n, _ in out))
for m in METRICS:
    if m in col:
        print(f"{m} [{units[col[m]]}]: " + " | ".join(r[col[m]] for _, r in out))

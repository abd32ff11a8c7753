"""ncu `--page raw --csv` exports -> one markdown table (duration, DRAM bytes and %, tensor pipe %, registers, dyn. smem).
usage: python tools/ncu_summary.py out.md raw1.csv raw2.csv ..."""
import csv, sys
COLS = [("duration", "gpu__time_duration.sum"), ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
        ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("dyn smem", "launch__shared_mem_per_block_dynamic"), ("grid", "launch__grid_size")]
out = ["| kernel | " + " | ".join(c for c, _ in COLS) + " |", "|---|" + "---|" * len(COLS)]
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        cells = []
        for _, m in COLS:
            if m in idx:
                u = units[idx[m]]
                v = r[idx[m]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        out.append(f"| `{name}` | " + " | ".join(cells) + " |")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))

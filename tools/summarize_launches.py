"""ncu --csv launch list (gpu__time_duration.sum) -> per-kernel totals and shares; usage: summarize_launches.py in.csv [out.csv]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hd = rows[h]
ki, vi = hd.index("Kernel Name"), hd.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > vi:
        k = r[ki].split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1000.0
tot = sum(t for _, t in agg.values())
lines = ["kernel,launches,total_us,share_of_step"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f'"{k}",{n},{t:.1f},{t / tot:.4f}')
lines.append(f"TOTAL,{sum(n for n, _ in agg.values())},{tot:.1f},1.0")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")

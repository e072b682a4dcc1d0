"""Summarise an ncu launch list (--csv --log-file ...) per kernel: total time, launches, share, DRAM bytes.
    python tools/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in data:
    if len(r) < len(hdr):
        continue
    k, m = r[ix["Kernel Name"]], r[ix["Metric Name"]]
    agg[k][m] += float(r[ix["Metric Value"]].replace(",", "")) * scale.get(r[ix["Metric Unit"]], 1.0)
    if m == "gpu__time_duration.sum":
        cnt[k] += 1
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
print("    total ms  launches    share  DRAM read GB  DRAM write GB  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    print(f"{a['gpu__time_duration.sum']:12.3f} {cnt[k]:9d} {100 * a['gpu__time_duration.sum'] / tot:7.3f}% "
          f"{a['dram__bytes_read.sum']:13.3f} {a['dram__bytes_write.sum']:14.3f}  {k[:100]}")

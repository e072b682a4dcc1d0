"""DRAM traffic of the observation kernel per 64-particle tile, from an ncu launch list that carried
dram__bytes_read.sum / dram__bytes_write.sum (tools/sessions/*.sh), appended to profiles/traffic_obs_kernel.json,
the file bench.py's roofline.traffic is computed from.
    python tools/ncu_traffic.py gpurun_out/launches_r02a.csv --N 20000 --d 3 --particles 37888 --source "..." """
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--N", type=int, required=True)
ap.add_argument("--d", type=int, required=True)
ap.add_argument("--particles", type=int, required=True)
ap.add_argument("--kernel", default="gp_predict_kernel<0")
ap.add_argument("--source", required=True)
ap.add_argument("--note", default="")
ap.add_argument("--no-cache", action="store_true")
ap.add_argument("--dense", action="store_true")
o = ap.parse_args()

rows = list(csv.reader(open(o.csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
per_launch = {}
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]) or o.kernel not in r[ix["Kernel Name"]]:
        continue
    m = r[ix["Metric Name"]]
    if m.startswith("dram__bytes_"):
        per_launch.setdefault(r[ix["ID"]], 0.0)
        per_launch[r[ix["ID"]]] += float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
vals = sorted(per_launch.values())
tiles = (o.particles + 63) // 64
rec = {"N": o.N, "d": o.d, "kstar_cache": not o.no_cache, "tri": not o.dense, "launches": len(vals),
       "tiles_per_launch": tiles, "dram_bytes_per_launch_median": vals[len(vals) // 2],
       "dram_bytes_per_tile": vals[len(vals) // 2] / tiles, "source": o.source, "note": o.note}
path = os.path.join(ROOT, "profiles", "traffic_obs_kernel.json")
data = json.load(open(path)) if os.path.exists(path) else {"captures": []}
data["captures"] = [c for c in data["captures"] if not (c["N"] == rec["N"] and c["d"] == rec["d"] and
                                                          c["kstar_cache"] == rec["kstar_cache"] and c["tri"] == rec["tri"])]
data["captures"].append(rec)
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(rec))

#!/bin/bash
# Round 2, GPU session AT: ncu launch list of the six small-cloud kernels (direct launches, P = 100, N = 2000).
mkdir -p gpurun_out
CMD="python tools/run_trials.py --trials 1 --no-graph"
timeout 300 $CMD > gpurun_out/plain_r02at.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gp_predict_kernel|small_pre_kernel|small_post_kernel|predict_finalize_kernel|kstar_fill_kernel" --csv --log-file gpurun_out/launches_small_r02at.csv $CMD > gpurun_out/ncu_small_at.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_small_r02at.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
seq = [(r[ix["Kernel Name"]], float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]) for r in rows[hi + 1:] if len(r) > ix["Metric Value"]]
print(len(seq), "launches")
tail = seq[-60:]
agg = collections.defaultdict(list)
for k, v, u in seq[len(seq) // 2:]:
    agg[k.split("(")[0][-60:]].append(v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0))
for k, v in agg.items():
    v.sort(); print(f"{k:62s} n={len(v):5d} median {v[len(v)//2]:8.2f} us")
PY

#!/bin/bash
# Round 2, GPU session U: per-chunk time of the fused cached launch vs the low-latency items (diagnostic build), SM clock.
mkdir -p gpurun_out
timeout 600 python tools/lowlat_timeline.py --classes 8 --frames 250 --fused > gpurun_out/timeline_fused_n20000.json 2> gpurun_out/timeline.err; tail -2 gpurun_out/timeline.err; cat gpurun_out/timeline_fused_n20000.json
timeout 600 python tools/lowlat_timeline.py --classes 8 --frames 250 > gpurun_out/timeline_n20000_u.json 2>> gpurun_out/timeline.err
timeout 600 python tools/lowlat_timeline.py --classes 8 --frames 250 --particles 64 > gpurun_out/timeline_n20000_p64.json 2>> gpurun_out/timeline.err
timeout 600 python tools/lowlat_timeline.py --fused > gpurun_out/timeline_fused_n2000.json 2>> gpurun_out/timeline.err; cat gpurun_out/timeline_fused_n2000.json
python - <<'PY'
import json
for f in ("gpurun_out/timeline_n20000_u.json", "gpurun_out/timeline_n20000_p64.json"):
    d = json.load(open(f)); print(d["workload"], d["sm_clock_mhz_after"], d["launch_ms_events"], d["last_epilogue_done_ns"], d["median"], d["items_real"])
PY

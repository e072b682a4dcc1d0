#!/bin/bash
# Round 2, GPU session AU: per-item timeline of the DYNAMICS GP's low-latency launch at the reference's operating point.
mkdir -p gpurun_out
timeout 300 python tools/lowlat_timeline.py --dynamics > gpurun_out/timeline_dyn_n2000.json 2> gpurun_out/timeline.err; tail -2 gpurun_out/timeline.err; cat gpurun_out/timeline_dyn_n2000.json
for sg in 8 16; do timeout 300 python tools/lowlat_timeline.py --dynamics --seg $sg > gpurun_out/timeline_dyn_n2000_seg$sg.json 2>> gpurun_out/timeline.err; python -c "
import json;d=json.load(open('gpurun_out/timeline_dyn_n2000_seg$sg.json'));print(d['workload'],d['launch_ms_events'],d['last_epilogue_done_ns'],d['median'],d['items_real'],d['items_empty'])"; done

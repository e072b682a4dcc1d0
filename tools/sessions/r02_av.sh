#!/bin/bash
# Round 2, GPU session AV: node timeline of one graph-replayed small-cloud frame (diagnostic build).
mkdir -p gpurun_out
timeout 300 python tools/lowlat_timeline.py --frame > gpurun_out/timeline_frame_n2000.json 2> gpurun_out/timeline.err; tail -3 gpurun_out/timeline.err
python -c "
import json;d=json.load(open('gpurun_out/timeline_frame_n2000.json'));print(d['workload']);print('kernels',d['kernels']);print('durations',d['durations_ns']);print('gaps',d['gaps_ns']);print('frame',d['frame_ns'])"

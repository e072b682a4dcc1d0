#!/bin/bash
# Round 2, GPU session S: per-item timeline of the low-latency observation launch (diagnostic build), achieved-error report.
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 300 python -m pytest tests/test_gpu_parity_golden.py -m gpu -q -k achieved 2>&1 | tail -3
cat gpurun_out/parity_achieved.jsonl
timeout 300 python tools/lowlat_timeline.py > gpurun_out/timeline_n2000.json 2> gpurun_out/timeline.err; tail -2 gpurun_out/timeline.err; cat gpurun_out/timeline_n2000.json
timeout 300 python tools/lowlat_timeline.py --seg 16 > gpurun_out/timeline_n2000_seg16.json 2>> gpurun_out/timeline.err; cat gpurun_out/timeline_n2000_seg16.json
timeout 300 python tools/lowlat_timeline.py --seg 6 > gpurun_out/timeline_n2000_seg6.json 2>> gpurun_out/timeline.err; cat gpurun_out/timeline_n2000_seg6.json
timeout 600 python tools/lowlat_timeline.py --classes 8 --frames 250 > gpurun_out/timeline_n20000.json 2>> gpurun_out/timeline.err; cat gpurun_out/timeline_n20000.json

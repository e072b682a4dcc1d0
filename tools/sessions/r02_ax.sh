#!/bin/bash
# Round 2, GPU session AX: smoke() and a short bench on the final tree.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --particles 37888 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_r02ax.json 2> gpurun_out/bench_small.err
python -c "import json;d=json.load(open('gpurun_out/bench_small_r02ax.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['launch_ms'],r['frac'],d['parity']['digest'][:12],d['clocks'])"

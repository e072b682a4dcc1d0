#!/bin/bash
# Round 2, GPU session AZ: TMA look-ahead of the fp64 kernel (chunks in flight): 3 (default) vs 4 vs 2, same stage count.
mkdir -p gpurun_out
for lib in "" _ahead4 _ahead2; do
  for i in 1 2; do
    GPMDM_LIBRARY=$PWD/gpmdm_b200/lib/libgpmdm_sm100a$lib.so timeout 600 python bench.py --particles 37888 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_ahead$lib.json 2> gpurun_out/bench_ahead.err
    python -c "import json;d=json.load(open('gpurun_out/bench_ahead$lib.json'));r=d['roofline'];print('lib$lib',d['value'],r['launch_ms'],r['frac'],d['parity']['digest'][:12])"
  done
done

#!/bin/bash
# Round 2, GPU session O: suite (dynamics K* cache), BASELINE config 2 bench, cfg4 tolerance check of both tensor-core variants.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02o.log
tail -4 gpurun_out/pytest_r02o.log
timeout 600 python bench.py --classes 2 --seqs-per-class 10 --frames 100 --particles 100000 --steps 5 --warmup 3 --cpu-sample 2000 > gpurun_out/bench_cfg2_fp64_r02.json 2> gpurun_out/bench_cfg2.err
python -c "import json;d=json.load(open('gpurun_out/bench_cfg2_fp64_r02.json'));r=d['roofline'];print('cfg2',d['value'],d['ms_per_step'],r['launch_ms'],r['frac'],r['executed_frac'],d['cpu_baseline'],{k:v for k,v in d['parity'].items() if k.endswith('equal') or k.endswith('_max')})"
tail -2 gpurun_out/bench_cfg2.err
timeout 900 python tools/cfg4_check.py > gpurun_out/cfg4_tolerance_check_r02.json 2> gpurun_out/cfg4_check.err
cat gpurun_out/cfg4_tolerance_check_r02.json; tail -2 gpurun_out/cfg4_check.err

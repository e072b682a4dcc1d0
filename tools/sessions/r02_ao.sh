#!/bin/bash
# Round 2, GPU session AO: final library after the triangle-aware factor products: suite, smoke(), headline bench, cfg2 bench.
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02ao.log
tail -4 gpurun_out/pytest_r02ao.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_full_r02ao.json 2> gpurun_out/bench_full_r02ao.err
python -c "import json;d=json.load(open('gpurun_out/bench_full_r02ao.json'));r=d['roofline'];p=d['parity'];print(d['value'],d['e2e']['value'],d['ms_per_step'],r['launch_ms'],r['frac'],r['executed_frac'],p['digest'],p['classes_equal'],p['ancestors_equal'],p['obs_var_err_of_prior_max'],p['ll_err_vs_bound_max'],d['cpu_baseline']['value'],d['clocks'])"
tail -2 gpurun_out/bench_full_r02ao.err
timeout 600 python bench.py --classes 2 --seqs-per-class 10 --frames 100 --particles 100000 --steps 5 --warmup 3 --cpu-sample 2000 > gpurun_out/bench_cfg2_fp64_r02ao.json 2> gpurun_out/bench_cfg2.err
python -c "import json;d=json.load(open('gpurun_out/bench_cfg2_fp64_r02ao.json'));r=d['roofline'];print('cfg2',d['value'],d['ms_per_step'],r['launch_ms'],r['frac'],d['parity']['digest'])"

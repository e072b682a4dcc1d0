#!/bin/bash
# Round 2, GPU session AI: final library: suite, smoke(), headline bench, reference arm (bounded).
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02ai.log
tail -4 gpurun_out/pytest_r02ai.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full_r02ai.json 2> gpurun_out/bench_full_r02ai.err
python -c "import json;d=json.load(open('gpurun_out/bench_full_r02ai.json'));r=d['roofline'];print(d['value'],d['e2e']['value'],d['ms_per_step'],r['launch_ms'],r['frac'],r['executed_frac'],d['parity']['digest'],d['parity']['classes_equal'],d['parity']['ancestors_equal'],d['cpu_baseline']['value'],d['clocks'])"
tail -2 gpurun_out/bench_full_r02ai.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r02ai.json 2> gpurun_out/bench_ref_r02ai.err
cut -c1-700 gpurun_out/bench_ref_r02ai.json; tail -2 gpurun_out/bench_ref_r02ai.err

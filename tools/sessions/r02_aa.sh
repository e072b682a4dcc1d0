#!/bin/bash
# Round 2, GPU session AA: Hadamard epilogue from the K* slices (CACHE instances): suite, headline bench (same digest?), cfg2 bench,
# small-cloud trials at N = 20 000.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02aa.log
tail -4 gpurun_out/pytest_r02aa.log
timeout 900 python bench.py > gpurun_out/bench_full_r02aa.json 2> gpurun_out/bench_full_r02aa.err
python -c "import json;d=json.load(open('gpurun_out/bench_full_r02aa.json'));r=d['roofline'];print(d['value'],d['e2e']['value'],d['ms_per_step'],r['launch_ms'],r['frac'],r['executed_frac'],d['parity']['digest'],d['parity']['classes_equal'],d['parity']['ancestors_equal'],d['clocks'])"
tail -2 gpurun_out/bench_full_r02aa.err
timeout 600 python bench.py --classes 2 --seqs-per-class 10 --frames 100 --particles 100000 --steps 5 --warmup 3 --cpu-sample 2000 > gpurun_out/bench_cfg2_fp64_r02aa.json 2> gpurun_out/bench_cfg2.err
python -c "import json;d=json.load(open('gpurun_out/bench_cfg2_fp64_r02aa.json'));r=d['roofline'];print('cfg2',d['value'],d['ms_per_step'],r['launch_ms'],r['frac'],r['executed_frac'],d['parity']['digest'])"
tail -2 gpurun_out/bench_cfg2.err
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 --batched > gpurun_out/trials_r02aa.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 6 --batched >> gpurun_out/trials_r02aa.jsonl 2>> gpurun_out/trials.err
python -c "
import json
for l in open('gpurun_out/trials_r02aa.jsonl'):
    j=json.loads(l); print(j['workload'][:58], '|', j['driver'][:24], round(j['seconds_per_frame']*1e3,4),'ms', round(j['fps']), j['frame_accuracy'])
"

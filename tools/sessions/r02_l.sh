#!/bin/bash
# Round 2, GPU session L: suite; stage kernels at a bandwidth-bound size; small-cloud trial driver (per-frame and batched);
# the headline bench line after the padding-chunk skip.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02l.log
tail -4 gpurun_out/pytest_r02l.log
timeout 300 python tools/k3_bench.py > gpurun_out/k3_stage_kernels_r02.json 2> gpurun_out/k3.err
python -c "import json;d=json.load(open('gpurun_out/k3_stage_kernels_r02.json'));print({k:(round(v['ms'],3),round(v['frac_of_peak_moved'],3)) for k,v in d['kernels'].items()})"
rm -f gpurun_out/trials_r02l.jsonl
python tools/run_trials.py --trials 6 >> gpurun_out/trials_r02l.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 6 --batched >> gpurun_out/trials_r02l.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 3 --particles 1000 --batched >> gpurun_out/trials_r02l.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 --batched >> gpurun_out/trials_r02l.jsonl 2>> gpurun_out/trials.err
python -c "
import json
for l in open('gpurun_out/trials_r02l.jsonl'):
    j=json.loads(l); print(j['workload'][:58], '|', j['driver'][:24], '|', j['step_path'], round(j['seconds_per_frame']*1e3,4),'ms', round(j['fps']), j['frame_accuracy'])
"
timeout 900 python bench.py > gpurun_out/bench_full_r02l.json 2> gpurun_out/bench_full_r02l.err
python -c "import json;d=json.load(open('gpurun_out/bench_full_r02l.json'));r=d['roofline'];print(d['value'],d['e2e']['value'],d['ms_per_step'],r['frac'],r['executed_frac'],d['cpu_baseline']['value'],d['parity'],d['clocks'])"
tail -2 gpurun_out/bench_full_r02l.err

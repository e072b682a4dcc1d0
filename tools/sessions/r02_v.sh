#!/bin/bash
# Round 2, GPU session V: small-cloud frame as six kernel nodes (no memset / copy nodes), shared K* slices only for large launches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_r02v.log
tail -6 gpurun_out/pytest_r02v.log
rm -f gpurun_out/trials_r02v.jsonl
python tools/run_trials.py --trials 6 >> gpurun_out/trials_r02v.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 6 --batched >> gpurun_out/trials_r02v.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 3 --particles 1000 --batched >> gpurun_out/trials_r02v.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 --batched >> gpurun_out/trials_r02v.jsonl 2>> gpurun_out/trials.err
python -c "
import json
for l in open('gpurun_out/trials_r02v.jsonl'):
    j=json.loads(l); print(j['workload'][:58], '|', j['driver'][:24], '|', j['step_path'], round(j['seconds_per_frame']*1e3,4),'ms', round(j['fps']), j['frame_accuracy'])
"
tail -2 gpurun_out/trials.err

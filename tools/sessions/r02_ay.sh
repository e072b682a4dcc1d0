#!/bin/bash
# Round 2, GPU sessions AW / AY: dynamics segment choice for one more ragged tile; post kernel with generic loads and its block partials in shared memory:
# small-cloud tests, frame timeline, trial driver, suite.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_small_cloud.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/lowlat_timeline.py --frame > gpurun_out/timeline_frame_n2000_ay.json 2> gpurun_out/timeline.err; tail -3 gpurun_out/timeline.err
python -c "
import json;d=json.load(open('gpurun_out/timeline_frame_n2000_ay.json'));print('durations',d['durations_ns']);print('gaps',d['gaps_ns']);print('frame',d['frame_ns'])"
rm -f gpurun_out/trials_r02ay.jsonl
python tools/run_trials.py --trials 6 >> gpurun_out/trials_r02ay.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 6 --batched >> gpurun_out/trials_r02ay.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 3 --particles 1000 --batched >> gpurun_out/trials_r02ay.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 --batched >> gpurun_out/trials_r02ay.jsonl 2>> gpurun_out/trials.err
python -c "
import json
for l in open('gpurun_out/trials_r02ay.jsonl'):
    j=json.loads(l); print(j['workload'][:58], '|', j['driver'][:24], round(j['seconds_per_frame']*1e3,4),'ms', round(j['fps']), j['frame_accuracy'])
"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02ay.log
tail -4 gpurun_out/pytest_r02ay.log

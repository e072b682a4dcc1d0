#!/bin/bash
# Round 2, GPU session T: low-latency launches with shared K* slices + trimmed ragged tiles: suite, timelines, trial driver.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_r02t.log
tail -6 gpurun_out/pytest_r02t.log
timeout 300 python tools/lowlat_timeline.py > gpurun_out/timeline_n2000_t.json 2> gpurun_out/timeline.err; tail -2 gpurun_out/timeline.err
timeout 600 python tools/lowlat_timeline.py --classes 8 --frames 250 > gpurun_out/timeline_n20000_t.json 2>> gpurun_out/timeline.err
python - <<'PY'
import json
for f in ("gpurun_out/timeline_n2000_t.json", "gpurun_out/timeline_n20000_t.json"):
    d = json.load(open(f)); print(d["workload"], d["launch_ms_events"], d["last_epilogue_done_ns"], d["median"], d["items_real"])
PY
rm -f gpurun_out/trials_r02t.jsonl
python tools/run_trials.py --trials 6 >> gpurun_out/trials_r02t.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 6 --batched >> gpurun_out/trials_r02t.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 3 --particles 1000 --batched >> gpurun_out/trials_r02t.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 --batched >> gpurun_out/trials_r02t.jsonl 2>> gpurun_out/trials.err
python -c "
import json
for l in open('gpurun_out/trials_r02t.jsonl'):
    j=json.loads(l); print(j['workload'][:58], '|', j['driver'][:24], '|', j['step_path'], round(j['seconds_per_frame']*1e3,4),'ms', round(j['fps']), j['frame_accuracy'])
"
tail -2 gpurun_out/trials.err

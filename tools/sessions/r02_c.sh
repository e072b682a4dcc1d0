#!/bin/bash
# Round 2, GPU session C: full GPU suite (all failures), per-kernel times of the small-cloud step (ncu launch list).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_r02c.log
tail -6 gpurun_out/pytest_r02c.log
CMD="python tools/run_trials.py --trials 1 --frames 12 --no-graph"
$CMD > gpurun_out/plain_r02c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_trials_r02c.csv $CMD > gpurun_out/ncu_trials.log 2>&1
tail -2 gpurun_out/ncu_trials.log
python tools/run_trials.py --trials 3 --particles 1000 >> gpurun_out/trials_p1000_r02.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 3 --particles 1000 --staged >> gpurun_out/trials_p1000_r02.jsonl 2>> gpurun_out/trials.err
cat gpurun_out/trials_p1000_r02.jsonl

#!/bin/bash
# Round 2, GPU session C: full GPU suite (all failures), the new fp16-split kernel under a timeout, per-kernel times of the
# small-cloud step (ncu launch list).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not f16x2" 2>&1 | tail -60 > gpurun_out/pytest_r02c.log
tail -6 gpurun_out/pytest_r02c.log
timeout 300 python -m pytest tests/test_gpu_tf32_variant.py -m gpu -q -x -k "f16x2" 2>&1 | tail -40 > gpurun_out/pytest_f16_r02c.log
tail -15 gpurun_out/pytest_f16_r02c.log
timeout 300 python bench.py --precision f16x2 --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_f16x2_r02c.json 2> gpurun_out/bench_f16.err
cat gpurun_out/bench_cfg3_f16x2_r02c.json; tail -3 gpurun_out/bench_f16.err
CMD="python tools/run_trials.py --trials 1 --frames 12 --no-graph"
$CMD > gpurun_out/plain_r02c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_trials_r02c.csv $CMD > gpurun_out/ncu_trials.log 2>&1
tail -2 gpurun_out/ncu_trials.log
python tools/run_trials.py --trials 3 --particles 1000 >> gpurun_out/trials_p1000_r02.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 3 --particles 1000 --staged >> gpurun_out/trials_p1000_r02.jsonl 2>> gpurun_out/trials.err
cat gpurun_out/trials_p1000_r02.jsonl

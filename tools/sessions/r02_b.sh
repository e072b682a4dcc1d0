#!/bin/bash
# Round 2, GPU session B: full GPU suite, cfg1 latency (trial driver), tf32 bench at cfg3 sizes, ncu full capture of the observation kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_r02b.log
tail -4 gpurun_out/pytest_r02b.log
for opt in "" "--no-graph" "--staged"; do
  python tools/run_trials.py --trials 6 $opt >> gpurun_out/trials_cfg1_r02.jsonl 2>> gpurun_out/trials.err
done
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 >> gpurun_out/trials_n20k_r02.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 --staged >> gpurun_out/trials_n20k_r02.jsonl 2>> gpurun_out/trials.err
cat gpurun_out/trials_cfg1_r02.jsonl gpurun_out/trials_n20k_r02.jsonl
python bench.py --precision tf32 --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_tf32_r02b.json 2> gpurun_out/bench_tf32.err
cat gpurun_out/bench_cfg3_tf32_r02b.json
CMD="python bench.py --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_r02b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gp_predict_kernel -s 1 -c 1 \
    -o gpurun_out/prof_obs_r02b $CMD > gpurun_out/ncu_full_b.log 2>&1
tail -3 gpurun_out/ncu_full_b.log
ls -la gpurun_out | tail -8

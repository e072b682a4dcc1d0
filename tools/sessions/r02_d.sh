#!/bin/bash
# Round 2, GPU session D: suite; small-cloud latency after the post-kernel / segment / graph-node changes; tensor-core variants
# at cfg3 and cfg4 sizes; ncu full capture of the fp16-split kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r02d.log
tail -5 gpurun_out/pytest_r02d.log
rm -f gpurun_out/trials_r02d.jsonl
python tools/run_trials.py --trials 6 >> gpurun_out/trials_r02d.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 6 --no-graph >> gpurun_out/trials_r02d.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 3 --particles 1000 >> gpurun_out/trials_r02d.jsonl 2>> gpurun_out/trials.err
python tools/run_trials.py --trials 2 --classes 8 --seqs-per-class 25 >> gpurun_out/trials_r02d.jsonl 2>> gpurun_out/trials.err
cut -c1-400 gpurun_out/trials_r02d.jsonl
for prec in f16x2 tf32; do
  timeout 300 python bench.py --precision $prec --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_${prec}_r02d.json 2> gpurun_out/bench_${prec}.err
  cat gpurun_out/bench_cfg3_${prec}_r02d.json
done
timeout 600 python bench.py --precision f16x2 --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_f16x2_P524288_r02d.json 2> gpurun_out/bench_cfg4.err
cat gpurun_out/bench_cfg4_f16x2_P524288_r02d.json; tail -2 gpurun_out/bench_cfg4.err
CMD="python bench.py --precision f16x2 --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_r02d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:observe_tf32_kernel -s 1 -c 1 \
    -o gpurun_out/prof_f16_r02d $CMD > gpurun_out/ncu_full_d.log 2>&1
tail -3 gpurun_out/ncu_full_d.log

#!/bin/bash
# Round 2, GPU session A: test-suite, factor precompute timing, bench (small + full), ncu launch list + full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_r02a.log
tail -5 gpurun_out/pytest_r02a.log
python tools/factor_bench.py --cfg 3 --old > gpurun_out/factor_bench_cfg3_r02.json 2> gpurun_out/factor_bench_cfg3.err
python tools/factor_bench.py --cfg 4 > gpurun_out/factor_bench_cfg4_r02.json 2> gpurun_out/factor_bench_cfg4.err
cat gpurun_out/factor_bench_cfg3_r02.json gpurun_out/factor_bench_cfg4_r02.json
python bench.py > gpurun_out/bench_full_r02a.json 2> gpurun_out/bench_full_r02a.err
cat gpurun_out/bench_full_r02a.json
CMD="python bench.py --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_r02a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/launches_r02a.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_r02a2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gp_predict_kernel -s 1 -c 1 \
    -o gpurun_out/prof_obs_r02a $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20

#!/bin/bash
# Round 2, GPU session AK: ncu launch list + full capture of the headline fp64 observation kernel on the FINAL library.
mkdir -p gpurun_out
CMD="python bench.py --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_r02ak.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/launches_r02ak.csv $CMD > gpurun_out/ncu_launches_ak.log 2>&1
cat gpurun_out/plain_r02ak.log | cut -c1-300
timeout 600 $CMD > gpurun_out/plain_r02ak2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gp_predict_kernel -s 1 -c 1 \
    -o gpurun_out/prof_obs_r02ak $CMD > gpurun_out/ncu_full_ak.log 2>&1
tail -2 gpurun_out/ncu_full_ak.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02ak.csv

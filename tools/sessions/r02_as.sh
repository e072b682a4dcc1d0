#!/bin/bash
# Round 2, GPU session AS: ncu --set full of the tf32 observation kernel at cfg4 sizes (C = 64, N = 50 176, d = 8) and at cfg3 sizes.
mkdir -p gpurun_out
CMD="python bench.py --precision tf32 --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_r02as.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:observe_tf32_kernel -s 1 -c 1 \
    -o gpurun_out/prof_tf32_cfg4_r02as $CMD > gpurun_out/ncu_full_as.log 2>&1
tail -2 gpurun_out/ncu_full_as.log
CMD="python bench.py --precision tf32 --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_r02as2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:observe_tf32_kernel -s 1 -c 1 \
    -o gpurun_out/prof_tf32_cfg3_r02as $CMD > gpurun_out/ncu_full_as2.log 2>&1
tail -2 gpurun_out/ncu_full_as2.log
ls -la gpurun_out/*r02as*.ncu-rep

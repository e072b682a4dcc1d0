#!/bin/bash
# Round 2, GPU session AF: ncu --set full of the fp16-split observation kernel at cfg4 sizes (C = 64, N = 50 176, d = 8).
mkdir -p gpurun_out
CMD="python bench.py --precision f16x2 --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_r02af.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:observe_tf32_kernel -s 1 -c 1 \
    -o gpurun_out/prof_f16_cfg4_r02af $CMD > gpurun_out/ncu_full_af.log 2>&1
tail -3 gpurun_out/ncu_full_af.log
ls -la gpurun_out/*.ncu-rep

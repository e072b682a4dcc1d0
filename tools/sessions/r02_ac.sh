#!/bin/bash
# Round 2, GPU session AC: the reference's published training run (500 Adam steps, N ~ 2000) on the GPU.
mkdir -p gpurun_out
timeout 900 python tools/train_bench.py > gpurun_out/train_bench_r02.json 2> gpurun_out/train_bench.err; tail -3 gpurun_out/train_bench.err; cat gpurun_out/train_bench_r02.json

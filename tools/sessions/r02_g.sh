#!/bin/bash
# Round 2, GPU session G: suite after the symmetric-walk gradient kernel; training-kernel bench (BASELINE config 5).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02g.log
tail -4 gpurun_out/pytest_r02g.log
timeout 600 python tools/cfg5_train_bench.py > gpurun_out/cfg5_training_kernels_r02.json 2> gpurun_out/cfg5.err
cat gpurun_out/cfg5_training_kernels_r02.json; tail -3 gpurun_out/cfg5.err

#!/bin/bash
# Round 2, GPU session AL: triangle-aware block recursions (triangular inverse, L^-T L^-1, panel products): training and factor
# tests, factor precompute timing at cfg3 / cfg4 sizes, the N = 20 000 training step, the published-shape training run.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_kernels.py tests/test_gpu_factor_precompute.py -m gpu -q 2>&1 | tail -4
timeout 600 python tools/factor_bench.py --cfg 3 > gpurun_out/factor_bench_cfg3_r02al.json 2> gpurun_out/factor_bench.err; cat gpurun_out/factor_bench_cfg3_r02al.json
timeout 900 python tools/factor_bench.py --cfg 4 > gpurun_out/factor_bench_cfg4_r02al.json 2>> gpurun_out/factor_bench.err; cat gpurun_out/factor_bench_cfg4_r02al.json
tail -2 gpurun_out/factor_bench.err
timeout 600 python tools/cfg5_train_bench.py > gpurun_out/cfg5_training_kernels_r02al.json 2> gpurun_out/cfg5.err; python -c "
import json;d=json.load(open('gpurun_out/cfg5_training_kernels_r02al.json'));print({k:v for k,v in d.items() if 'ms' in k})"
tail -2 gpurun_out/cfg5.err
timeout 600 python tools/train_bench.py > gpurun_out/train_bench_r02al.json 2> gpurun_out/train_bench.err; python -c "
import json;d=json.load(open('gpurun_out/train_bench_r02al.json'));print(d['train_wall_s'], d['ms_per_adam_step'], d['loss_last'])"

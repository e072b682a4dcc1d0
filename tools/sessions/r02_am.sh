#!/bin/bash
# Round 2, GPU session AM: where the N = 20 000 training step's 358 ms go (torch.profiler kernel table).
mkdir -p gpurun_out
timeout 600 python tools/cfg5_train_bench.py --profile > gpurun_out/cfg5_profile.json 2> gpurun_out/cfg5_profile.err
grep -v "^$" gpurun_out/cfg5_profile.err | cut -c1-200 | tail -40

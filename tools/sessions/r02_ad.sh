#!/bin/bash
# Round 2, GPU session AD: long training trajectories against the reference's recorded runs.
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 900 python -m pytest tests/test_gpu_training_kernels.py -m gpu -q -k "trajectory" 2>&1 | tail -15
cat gpurun_out/parity_achieved.jsonl

#!/bin/bash
# Round 2, GPU session AB: the cfg1-scale reference fixture (product's own factors) + achieved-error report; suite.
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 600 python -m pytest tests/test_gpu_parity_golden.py -m gpu -q -k "achieved or cfg1" 2>&1 | tail -15
cat gpurun_out/parity_achieved.jsonl | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02ab.log
tail -4 gpurun_out/pytest_r02ab.log

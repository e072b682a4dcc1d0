#!/bin/bash
# Round 2, GPU session Y: suite after the guard-test extension (both K* modes, NaN-poisoned workspace) and the ABI constant.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02y.log
tail -6 gpurun_out/pytest_r02y.log

#!/bin/bash
# Round 2, GPU session AE: programmatic dependent launch for the six-kernel small-cloud frame: small-cloud tests first (under a
# timeout), trial driver with and without PDL, then the suite.
# (The PDL code this session measured lives in commit 2ab1e80 only: it was 8 % slower and was reverted; GPMDM_PDL has no effect now.)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_small_cloud.py -m gpu -q -x 2>&1 | tail -5
rm -f gpurun_out/trials_r02ae.jsonl
for pdl in 1 0; do
  GPMDM_PDL=$pdl timeout 300 python tools/run_trials.py --trials 6 >> gpurun_out/trials_r02ae.jsonl 2>> gpurun_out/trials.err
  GPMDM_PDL=$pdl timeout 300 python tools/run_trials.py --trials 6 --batched >> gpurun_out/trials_r02ae.jsonl 2>> gpurun_out/trials.err
  GPMDM_PDL=$pdl timeout 300 python tools/run_trials.py --trials 6 --no-graph >> gpurun_out/trials_r02ae.jsonl 2>> gpurun_out/trials.err
done
GPMDM_PDL=1 timeout 300 python tools/run_trials.py --trials 3 --particles 1000 --batched >> gpurun_out/trials_r02ae.jsonl 2>> gpurun_out/trials.err
python -c "
import json
for l in open('gpurun_out/trials_r02ae.jsonl'):
    j=json.loads(l); print(j['workload'][:58], '|', j['driver'][:24], '|', j['step_path'][:40], round(j['seconds_per_frame']*1e3,4),'ms', round(j['fps']), j['frame_accuracy'])
"
tail -2 gpurun_out/trials.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02ae.log
tail -4 gpurun_out/pytest_r02ae.log

#!/bin/bash
# Round 2, 8-GPU session AQ (final library): bench.py as the driver launches it at N = 8 (default workload, P = 1 048 576 sharded).
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_cfg3_fp64_8gpu_r02aq.json 2> gpurun_out/bench_8gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_cfg3_fp64_8gpu_r02aq.json") if l.startswith("{")][-1])
    print(d["n_gpus"], round(d["value"]), d["ms_per_step"], d["roofline"]["executed_frac"], (d.get("parity") or {}).get("digest"), d["clocks"])
except Exception as e:
    print("unreadable:", e)
PY
tail -3 gpurun_out/bench_8gpu.err

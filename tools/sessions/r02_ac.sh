#!/bin/bash
# Round 2, GPU session AC: the reference's published training run (500 Adam steps, N ~ 2000) on the GPU, at two synthetic noise levels.
mkdir -p gpurun_out
timeout 900 python tools/train_bench.py > gpurun_out/train_bench_r02.json 2> gpurun_out/train_bench.err; tail -3 gpurun_out/train_bench.err
timeout 900 python tools/train_bench.py --noise 0.02 > gpurun_out/train_bench_noise002_r02.json 2>> gpurun_out/train_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/train_bench_r02.json", "gpurun_out/train_bench_noise002_r02.json"):
    d = json.load(open(f)); print(d["workload"][-60:], round(d["train_wall_s"], 2), "s", {k: (v["frame_accuracy"], v["variance_faults"], round(v["sigma_n_y"], 4), round(v["sigma_n_x"], 4)) for k, v in d["checkpoints"].items()})
PY

#!/bin/bash
# Round 2, 2-GPU session Z (final library): multi-GPU tests; bench.py at N = 1 and N = 2 on the same total particle count (same digest).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_multi_2gpu_r02z.log; cat gpurun_out/pytest_multi_2gpu_r02z.log
timeout 600 python bench.py --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_P262144_1gpu_r02z.json 2> gpurun_out/bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_P262144_2gpu_r02z.json 2> gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_P262144_1gpu_r02z.json", "gpurun_out/bench_P262144_2gpu_r02z.json"):
    try:
        lines = [l for l in open(f) if l.startswith("{")]
        d = json.loads(lines[-1])
        p = d.get("parity") or {}
        print(f, len(lines), "line(s)", d["n_gpus"], round(d["value"]), d["scaling"], p.get("digest"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/bench_2gpu.err

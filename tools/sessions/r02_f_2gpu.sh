#!/bin/bash
# Round 2, 2-GPU session: bit-equality of the sharded filter (pytest + state digests of bench.py at N = 1 and N = 2).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_multi_r02.log
cat gpurun_out/pytest_multi_r02.log
timeout 600 python bench.py --particles 262144 --steps 2 --warmup 3 --cpu-sample 1024 > gpurun_out/bench_P262144_1gpu_r02.json 2> gpurun_out/bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --particles 262144 --steps 2 --warmup 3 > gpurun_out/bench_P262144_2gpu_r02.json 2> gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_P262144_1gpu_r02.json", "gpurun_out/bench_P262144_2gpu_r02.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["n_gpus"], d["value"], d["parity"].get("digest"), {k: v for k, v in d["parity"].items() if k.endswith("equal") or "err" in k and not k.endswith("bound")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/bench_2gpu.err

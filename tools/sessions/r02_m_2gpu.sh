#!/bin/bash
# Round 2, 2-GPU session M: bench.py end to end at N = 1 and N = 2 (same digest = bit-identical sharded run), reference arm under torchrun.
mkdir -p gpurun_out
timeout 600 python bench.py --particles 262144 --steps 2 --warmup 3 --cpu-sample 1024 > gpurun_out/bench_P262144_1gpu_r02m.json 2> gpurun_out/bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --particles 262144 --steps 2 --warmup 3 > gpurun_out/bench_P262144_2gpu_r02m.json 2> gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 512 > gpurun_out/bench_ref_2gpu_r02m.json 2> gpurun_out/bench_ref.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_P262144_1gpu_r02m.json", "gpurun_out/bench_P262144_2gpu_r02m.json", "gpurun_out/bench_ref_2gpu_r02m.json"):
    try:
        lines = [l for l in open(f) if l.startswith("{")]
        d = json.loads(lines[-1])
        p = d.get("parity") or {}
        print(f, len(lines), "line(s)", d.get("impl", "ours"), d["n_gpus"], round(d["value"]), p.get("digest"), {k: v for k, v in p.items() if k.endswith("equal") or k.endswith("_max")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/bench_2gpu.err

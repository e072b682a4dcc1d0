#!/bin/bash
# Round 2, 2-GPU session AJ: the fp16-split variant (tensor-core observation + dynamics variance) at 1 and 2 GPUs: same digest?
mkdir -p gpurun_out
timeout 600 python bench.py --precision f16x2 --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f16x2_P262144_1gpu_r02aj.json 2> gpurun_out/bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --precision f16x2 --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f16x2_P262144_2gpu_r02aj.json 2> gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_f16x2_P262144_1gpu_r02aj.json", "gpurun_out/bench_f16x2_P262144_2gpu_r02aj.json"):
    try:
        lines = [l for l in open(f) if l.startswith("{")]
        d = json.loads(lines[-1])
        print(f, len(lines), "line(s)", d["n_gpus"], round(d["value"]), (d.get("parity") or {}).get("digest"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/bench_2gpu.err

#!/bin/bash
# Round 2, 2-GPU session AP (final library): multi-GPU tests; fp64 and fp16-split benches at 1 and 2 GPUs on the same total particle count.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_multi_2gpu_r02ap.log; cat gpurun_out/pytest_multi_2gpu_r02ap.log
for prec in fp64 f16x2; do
  timeout 600 python bench.py --precision $prec --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${prec}_P262144_1gpu_r02ap.json 2> gpurun_out/bench_1gpu.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --precision $prec --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${prec}_P262144_2gpu_r02ap.json 2> gpurun_out/bench_2gpu.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_P262144_?gpu_r02ap.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["n_gpus"], round(d["value"]), (d.get("parity") or {}).get("digest"))
    except Exception as e:
        print(f, "unreadable:", e)
PY

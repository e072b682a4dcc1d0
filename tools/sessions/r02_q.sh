#!/bin/bash
# Round 2, GPU session Q: tensor-core dynamics variance with the low-rank (linear kernel) part finished in fp64.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32_variant.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r02q_tc.log
tail -25 gpurun_out/pytest_r02q_tc.log
timeout 300 python tools/dynvar_check.py > gpurun_out/dynvar_check_r02.json 2> gpurun_out/dynvar.err; cat gpurun_out/dynvar_check_r02.json; tail -3 gpurun_out/dynvar.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r02q.log
tail -4 gpurun_out/pytest_r02q.log
for prec in f16x2 tf32; do
  timeout 600 python bench.py --precision $prec --particles 262144 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_${prec}_dyn_r02.json 2> gpurun_out/bench_${prec}.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg3_${prec}_dyn_r02.json'));r=d['roofline'];print('$prec',d['value'],d['ms_per_step'],r.get('launch_ms'),r.get('frac'),d.get('parity'))"
  tail -2 gpurun_out/bench_${prec}.err
done

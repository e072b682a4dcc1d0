#!/bin/bash
# Round 2, GPU session AG: 16 generator warps for d >= 5 in the tensor-core kernel: variant tests (under a timeout), cfg4 benches.
# (The 16-generator-warp kernel this session measured lives in commit a784b65 only: cfg4 f16x2 145.4 k -> 135.4 k/s; reverted.)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32_variant.py tests/test_gpu_benchmark_configs.py -m gpu -q -x 2>&1 | tail -5
for prec in f16x2 tf32; do
  timeout 900 python bench.py --precision $prec --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${prec}_g16_P524288_r02.json 2> gpurun_out/bench_cfg4.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg4_${prec}_g16_P524288_r02.json'));r=d['roofline'];print('cfg4 $prec',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
  tail -2 gpurun_out/bench_cfg4.err
done
timeout 900 python tools/cfg4_check.py > gpurun_out/cfg4_tolerance_check_r02ag.json 2> gpurun_out/cfg4_check.err
cat gpurun_out/cfg4_tolerance_check_r02ag.json; tail -2 gpurun_out/cfg4_check.err

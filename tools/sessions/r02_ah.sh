#!/bin/bash
# Round 2, GPU session AH: generators with two particle rows per thread (half the record loads): variant tests under a
# timeout, then the variant benches at cfg3 and cfg4 sizes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32_variant.py tests/test_gpu_benchmark_configs.py -m gpu -q -x 2>&1 | tail -5
for prec in f16x2 tf32; do
  timeout 600 python bench.py --precision $prec --particles 262144 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_${prec}_r2_r02.json 2> gpurun_out/bench_${prec}.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg3_${prec}_r2_r02.json'));r=d['roofline'];print('cfg3 $prec',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
  tail -2 gpurun_out/bench_${prec}.err
done
for prec in f16x2 tf32; do
  timeout 900 python bench.py --precision $prec --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${prec}_r2_P524288_r02.json 2> gpurun_out/bench_cfg4.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg4_${prec}_r2_P524288_r02.json'));r=d['roofline'];print('cfg4 $prec',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
  tail -2 gpurun_out/bench_cfg4.err
done

#!/bin/bash
# Round 2, GPU session K: cfg4 sizes, tensor-core variants with and without CTA pairs.
mkdir -p gpurun_out
for cl in 0 1; do for prec in f16x2 tf32; do
  GPMDM_TC_CLUSTER=$cl timeout 600 python bench.py --precision $prec --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${prec}_cl${cl}_P524288_r02k.json 2> gpurun_out/bench_cfg4.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg4_${prec}_cl${cl}_P524288_r02k.json'));r=d['roofline'];print('cfg4 cluster=$cl $prec',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
done; done

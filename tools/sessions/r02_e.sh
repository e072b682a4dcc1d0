#!/bin/bash
# Round 2, GPU session E: split timing of the tensor-core variants (tensor kernel vs fp64 mean tile), cfg4 tf32 for reference,
# ncu full capture of the fp64 kernels inside the f16x2 step (dynamics + mean tile).
mkdir -p gpurun_out
for prec in f16x2 tf32; do
  timeout 300 python bench.py --precision $prec --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_${prec}_r02e.json 2> gpurun_out/bench_${prec}.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg3_${prec}_r02e.json'));r=d['roofline'];print('$prec',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
done
timeout 600 python bench.py --precision tf32 --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_tf32_P524288_r02e.json 2> gpurun_out/bench_cfg4.err
python -c "import json;d=json.load(open('gpurun_out/bench_cfg4_tf32_P524288_r02e.json'));r=d['roofline'];print('cfg4 tf32',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
CMD="python bench.py --precision f16x2 --particles 37888 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_r02e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gp_predict_kernel -s 2 -c 2 \
    -o gpurun_out/prof_fp64parts_r02e $CMD > gpurun_out/ncu_full_e.log 2>&1
tail -3 gpurun_out/ncu_full_e.log

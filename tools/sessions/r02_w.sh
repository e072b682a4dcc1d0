#!/bin/bash
# Round 2, GPU session W: cfg4 sizes with the tensor-core dynamics variance; dynamics-variance error at cfg4 sizes; the
# headline fp64 bench line on the final library.
mkdir -p gpurun_out
for prec in f16x2 tf32; do
  timeout 900 python bench.py --precision $prec --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${prec}_dyn_P524288_r02.json 2> gpurun_out/bench_cfg4.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg4_${prec}_dyn_P524288_r02.json'));r=d['roofline'];print('cfg4 $prec',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
  tail -2 gpurun_out/bench_cfg4.err
done
timeout 600 python tools/dynvar_check.py --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --sample 4096 > gpurun_out/dynvar_check_cfg4_r02.json 2> gpurun_out/dynvar.err; cat gpurun_out/dynvar_check_cfg4_r02.json; tail -2 gpurun_out/dynvar.err
timeout 900 python bench.py > gpurun_out/bench_full_r02w.json 2> gpurun_out/bench_full_r02w.err
python -c "import json;d=json.load(open('gpurun_out/bench_full_r02w.json'));r=d['roofline'];print(d['value'],d['e2e']['value'],d['ms_per_step'],r['frac'],r['executed_frac'],d['cpu_baseline']['value'],d['parity'],d['clocks'])"
tail -2 gpurun_out/bench_full_r02w.err

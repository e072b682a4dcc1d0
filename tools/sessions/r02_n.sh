#!/bin/bash
# Round 2, GPU session N: ring-depth experiment for the single-CTA tensor-core kernel (A stages x B stages), f16x2.
mkdir -p gpurun_out
for lib in default a6b3 a4b4 a5b3; do
  if [ $lib = default ]; then unset GPMDM_LIBRARY; else export GPMDM_LIBRARY=$PWD/gpmdm_b200/lib/libgpmdm_tc_$lib.so; fi
  timeout 300 python bench.py --precision f16x2 --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n_cfg3_$lib.json 2> gpurun_out/n.err
  python -c "import json;d=json.load(open('gpurun_out/n_cfg3_$lib.json'));r=d['roofline'];print('$lib cfg3 f16x2',round(d['value']),r['launch_ms'],round(r['frac'],3),d['clocks']['sm_mhz'])"
  timeout 600 python bench.py --precision f16x2 --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/n_cfg4_$lib.json 2> gpurun_out/n.err
  python -c "import json;d=json.load(open('gpurun_out/n_cfg4_$lib.json'));r=d['roofline'];print('$lib cfg4 f16x2',round(d['value']),r['launch_ms'],round(r['frac'],3),d['clocks']['sm_mhz'])"
done

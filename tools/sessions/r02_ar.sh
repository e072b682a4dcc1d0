#!/bin/bash
# Round 2, GPU session AR: cross-chunk B-fragment prefetch in the fp64 k loop: bit-identity tests, then the kernel's launch time.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_guards.py tests/test_gpu_benchmark_configs.py tests/test_gpu_parity_golden.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --particles 37888 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_small_r02ar_$i.json 2> gpurun_out/bench_small.err
python -c "import json;d=json.load(open('gpurun_out/bench_small_r02ar_$i.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['launch_ms'],r['frac'],d['parity']['digest'][:12])"
done
timeout 600 python bench.py --classes 2 --seqs-per-class 10 --frames 100 --particles 100000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_r02ar.json 2> gpurun_out/bench_cfg2.err
python -c "import json;d=json.load(open('gpurun_out/bench_cfg2_r02ar.json'));r=d['roofline'];print('cfg2',d['value'],d['ms_per_step'],r['launch_ms'],d['parity']['digest'][:12])"

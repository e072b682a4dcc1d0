#!/bin/bash
# Round 2, GPU session J (records staged in shared memory): the cta_group::2 variant of the tensor-core kernel -- tests under a timeout (bounded waits trap
# instead of hanging), then the tensor-core benches at cfg3 / cfg4 sizes with and without clusters.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tf32_variant.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_tc_r02j.log
tail -12 gpurun_out/pytest_tc_r02j.log
if grep -q "passed" gpurun_out/pytest_tc_r02j.log && ! grep -q "failed\|error" gpurun_out/pytest_tc_r02j.log; then
  for cl in 1 0; do for prec in f16x2 tf32; do
    GPMDM_TC_CLUSTER=$cl timeout 300 python bench.py --precision $prec --particles 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_${prec}_cl${cl}_r02j.json 2> gpurun_out/bench_h.err
    python -c "import json;d=json.load(open('gpurun_out/bench_cfg3_${prec}_cl${cl}_r02j.json'));r=d['roofline'];print('cluster=$cl $prec',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
  done; done
  timeout 600 python bench.py --precision f16x2 --classes 64 --seqs-per-class 8 --frames 98 --latent 8 --particles 524288 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_f16x2_P524288_r02j.json 2> gpurun_out/bench_cfg4.err
  python -c "import json;d=json.load(open('gpurun_out/bench_cfg4_f16x2_P524288_r02j.json'));r=d['roofline'];print('cfg4 f16x2 cluster',d['value'],d['ms_per_step'],r['launch_ms'],r['fp64_mean_tile_ms'],r['achieved'],r['peak'],r['frac'],d['clocks'])"
fi

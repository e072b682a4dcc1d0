#!/bin/bash
# Round 2, GPU session AN: leaf sizes of the block recursions (K^-1 from the Cholesky factor; in-place triangular inverse).
mkdir -p gpurun_out
timeout 900 python tools/leaf_sweep.py > gpurun_out/leaf_sweep_r02.json 2> gpurun_out/leaf_sweep.err; tail -2 gpurun_out/leaf_sweep.err
python -c "
import json;d=json.load(open('gpurun_out/leaf_sweep_r02.json'))
for k in ('spd_inverse_from_cholesky_ms','tril_inverse_inplace_ms'):
    print(k); [print('  ',a,b) for a,b in d[k].items()]
print(d.get('max_rel_diff_between_settings'))"
